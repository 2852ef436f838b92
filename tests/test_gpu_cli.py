"""GPU: the featExtract CLI (3d_sift_cuda_b200/featExtract, the kept host code + the CUDA library)
against the reference's own CLI built from its sources (oracle/_ref/featExtract_ref, CPU path):
the feature files must be byte-identical."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "3d_sift_cuda_b200", "featExtract")
REF = os.path.join(ROOT, "oracle", "_ref", "featExtract_ref")


def run(exe, args, cwd):
    r = subprocess.run([exe] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-2000:]


CASES = [
    ("plain", [], dict()),
    ("double", ["-2+"], dict()),
    ("halve", ["-2-"], dict()),
    ("world_iso", ["-w"], dict(pixdim=(1.5, 1.5, 1.5), qoffset=(-40.0, 12.5, 7.0), quatern=(0.0, 0.0, 0.0))),
    ("world_rot_aniso", ["-w"], dict(pixdim=(1.0, 1.0, 2.0), qoffset=(3.0, -2.0, 1.0), quatern=(0.1, 0.2, 0.3))),
    # typed NIfTI voxels: our CLI sends them to the device as they are (cast there), the reference casts on the host
    ("int16", [], dict(dtype=np.int16)),
    ("uint8_double", ["-2+"], dict(dtype=np.uint8)),
    ("float64_world_aniso", ["-w"], dict(dtype=np.float64, pixdim=(1.0, 2.0, 1.0), qoffset=(1.0, 2.0, 3.0))),
]


@pytest.mark.parametrize("name,flags,hdr", CASES, ids=[c[0] for c in CASES])
def test_cli_matches_reference_cli(pkg, engine, tmp_path, name, flags, hdr):
    if not (os.path.exists(OURS) and os.path.exists(REF)):
        pytest.skip("CLI binaries not built")
    shape = (72, 64, 80) if name == "halve" else (40, 44, 36) if "double" in name else (56, 60, 52)
    vol = pkg.phantom.blob_phantom(shape, 31, 45)
    if hdr.get("dtype") is np.int16:
        vol = vol * 40.0          # use the integer range
    elif hdr.get("dtype") is np.uint8:
        vol = np.clip(vol * 2.0, 0, 255)
    nii = str(tmp_path / "in.nii")
    pkg.phantom.write_nifti(nii, vol, **hdr)
    run(REF, flags + [nii, "ref.key"], str(tmp_path))
    run(OURS, flags + [nii, "ours.key"], str(tmp_path))
    a = open(tmp_path / "ref.key", "rb").read()
    b = open(tmp_path / "ours.key", "rb").read()
    assert a.count(b"\n") > 8, "reference produced no features"
    assert a == b


def test_cli_raw_input_and_brief_flags(pkg, oracle, engine, tmp_path):
    if not os.path.exists(OURS):
        pytest.skip("CLI not built")
    vol = pkg.phantom.blob_phantom((64, 64, 64), 0, 60)
    raw = str(tmp_path / "in.f32")
    vol.tofile(raw)
    for flag, desc in (([], 0), (["-b"], 1), (["-br"], 2), (["-bn"], 3)):
        run(OURS, flag + ["-r", "64", "64", "64", raw, "o.key"], str(tmp_path))
        want = str(tmp_path / "w.key")
        pkg.api.write_features_text(want, oracle.extract(vol, 0, desc)["features"], (64, 64, 64))
        assert open(tmp_path / "o.key", "rb").read() == open(want, "rb").read(), flag


def test_cli_rejects_bad_arguments(tmp_path):
    if not os.path.exists(OURS):
        pytest.skip("CLI not built")
    r = subprocess.run([OURS, "-q", "a", "b"], stdout=subprocess.PIPE)
    assert r.returncode != 0 and b"unknown command line argument" in r.stdout
    r = subprocess.run([OURS, str(tmp_path / "missing.nii"), "o.key"], stdout=subprocess.PIPE)
    assert r.returncode != 0 and b"could not read input file" in r.stdout


REF_S3D = os.path.join(ROOT, "oracle", "_ref", "featExtract_ref_s3d")


@pytest.mark.parametrize("flags", [[], ["-2+"], ["-2-"]], ids=["plain", "double", "halve"])
def test_stage_level_drop_in_under_the_reference_pipeline(pkg, engine, tmp_path, flags):
    """INTEGRATION.md level B: the reference's OWN pipeline (unmodified src_common + featExtract.cpp) run with -d0,
    its four CUDA launchers replaced by oracle/shim/s3d_launchers.cpp on top of s3d_blur3d / s3d_dog /
    s3d_subsample2 / s3d_detect, must write the same bytes as the reference's CPU path."""
    if not (os.path.exists(REF_S3D) and os.path.exists(REF)):
        pytest.skip("reference binaries not built")
    shape = (72, 64, 80) if flags == ["-2-"] else (40, 44, 36) if flags == ["-2+"] else (56, 60, 52)
    vol = pkg.phantom.blob_phantom(shape, 31, 45)
    nii = str(tmp_path / "in.nii")
    pkg.phantom.write_nifti(nii, vol)
    run(REF, flags + [nii, "cpu.key"], str(tmp_path))
    run(REF_S3D, flags + ["-d0", nii, "s3d.key"], str(tmp_path))
    a = open(tmp_path / "cpu.key", "rb").read()
    b = open(tmp_path / "s3d.key", "rb").read()
    assert a.count(b"\n") > 8, "reference produced no features"
    assert a == b


def test_cli_slab_mode_over_listed_devices(pkg, engine, tmp_path):
    """featExtract -d0,0: ONE volume split into z slabs over the listed devices (here the same GPU twice) through
    s3d_multi_extract_slab must write the reference CLI's bytes."""
    if not (os.path.exists(OURS) and os.path.exists(REF)):
        pytest.skip("CLI binaries not built")
    vol = pkg.phantom.blob_phantom((48, 44, 230), 7, 120)
    nii = str(tmp_path / "in.nii")
    pkg.phantom.write_nifti(nii, vol)
    run(REF, [nii, "ref.key"], str(tmp_path))
    run(OURS, ["-d0,0", nii, "slab.key"], str(tmp_path))
    a = open(tmp_path / "ref.key", "rb").read()
    assert a.count(b"\n") > 8
    assert a == open(tmp_path / "slab.key", "rb").read()


def test_cli_list_mode_shards_volumes(pkg, engine, tmp_path):
    """featExtract -l list [-dA,B]: every "<input> <output>" line of the list is extracted (runs of equal shapes are
    sharded over the devices through s3d_multi_batch_extract); each output equals the reference CLI's."""
    if not (os.path.exists(OURS) and os.path.exists(REF)):
        pytest.skip("CLI binaries not built")
    shapes = [(56, 60, 52)] * 3 + [(40, 44, 36)] * 2
    lines = []
    for k, shp in enumerate(shapes):
        nii = str(tmp_path / ("in%d.nii" % k))
        pkg.phantom.write_nifti(nii, pkg.phantom.blob_phantom(shp, 40 + k, 45), pixdim=(1.0, 1.0, 1.0 if k else 2.0))
        run(REF, ["-w", nii, "ref%d.key" % k], str(tmp_path))
        lines.append("%s %s" % (nii, str(tmp_path / ("out%d.key" % k))))
    (tmp_path / "list.txt").write_text("\n".join(lines) + "\n")
    run(OURS, ["-w", "-d0,0", "-l", "list.txt"], str(tmp_path))
    for k in range(len(shapes)):
        assert open(tmp_path / ("ref%d.key" % k), "rb").read() == open(tmp_path / ("out%d.key" % k), "rb").read(), k


def test_cli_guesses_raw_dimensions(pkg, oracle, engine, tmp_path):
    """-r without dimensions: they are recovered from the signal (FeatureIO.cpp:3006-3227 idea, deterministic)."""
    if not os.path.exists(OURS):
        pytest.skip("CLI not built")
    vol = pkg.phantom.blob_phantom((72, 56, 40), 3, 60)
    raw = str(tmp_path / "in.f32")
    vol.tofile(raw)
    run(OURS, ["-r", raw, "o.key"], str(tmp_path))
    want = str(tmp_path / "w.key")
    pkg.api.write_features_text(want, oracle.extract(vol, 0, 0)["features"], (72, 56, 40))
    assert open(tmp_path / "o.key", "rb").read() == open(want, "rb").read()
