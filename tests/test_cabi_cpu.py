"""CPU: the C-ABI library loads, exports every symbol include/s3d.h declares, fails loudly without a
GPU, and its host-side pieces (taps, feature-file writer) match the reference.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "s3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(s3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.load_library()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert sorted(pkg.api.EXPORTS) == names, "api.EXPORTS and include/s3d.h disagree"


def test_feature_record_layout(pkg):
    assert pkg.FEATURE_DTYPE.itemsize == 324     # Feature3DInfo: uint + 4 + 9 + 3 + 64 floats
    assert pkg.CAND_DTYPE.itemsize == 16 and pkg.KEYPOINT_DTYPE.itemsize == 40


def test_no_gpu_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.S3DError):
        pkg.Engine(0)


def test_product_does_not_reference_the_oracle():
    """The product path must never import, link or execute anything under oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "3d_sift_cuda_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oracle|liboracle|libref3dsift|s3o_|ref_extract", src):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


@pytest.mark.parametrize("sigma", [0.5, 0.95, 1.2263, 1.5199, 3.09, 4.0])
def test_host_taps_match_oracle(pkg, oracle, sigma):
    assert pkg.gaussian_taps(sigma).tobytes() == oracle.taps(sigma).tobytes()


def test_text_writer_matches_reference_writer(pkg, tmp_path):
    gold = np.load(os.path.join(HERE, "golden", "ref_golden.npz"))
    feats = gold["blob64_features"]
    out = str(tmp_path / "f.key")
    pkg.api.write_features_text(out, feats, (64, 64, 64))
    assert open(out, "rb").read() == open(os.path.join(HERE, "golden", "blob64_ref.key"), "rb").read()


def test_binary_writer_matches_reference_writer(pkg, tmp_path):
    """msFeature3DVectorOutputBin (MultiScale.h:228-303): byte-identical files, with and without the eigenvalue filter."""
    gold = np.load(os.path.join(HERE, "golden", "ref_golden.npz"))
    feats = gold["blob64_features"]
    for name, thres in (("blob64_ref.bin", -1.0), ("blob64_ref_eig140.bin", 140.0)):
        out = str(tmp_path / name)
        pkg.api.write_features_bin(out, feats, thres)
        assert open(out, "rb").read() == open(os.path.join(HERE, "golden", name), "rb").read(), name


def test_text_reader_matches_reference_reader(pkg, tmp_path):
    """msFeature3DVectorInputText (MultiScale.h:305-384) on the reference's own .key file, and a write -> read round trip."""
    want = np.load(os.path.join(HERE, "golden", "blob64_key_readback.npy"))
    got = pkg.api.read_features_text(os.path.join(HERE, "golden", "blob64_ref.key"))
    assert len(got) == len(want) > 0 and got.tobytes() == want.tobytes()
    out = str(tmp_path / "rt.key")
    pkg.api.write_features_text(out, got, (64, 64, 64))
    again = pkg.api.read_features_text(out)
    assert again.tobytes() == got.tobytes()
    with pytest.raises(pkg.S3DError):
        pkg.api.read_features_text(str(tmp_path / "missing.key"))
    bad = tmp_path / "bad.key"
    bad.write_text("# comment\nFeatures: 0\n")
    with pytest.raises(pkg.S3DError):
        pkg.api.read_features_text(str(bad))


def test_feature_io_against_live_reference(pkg, reference, tmp_path):
    """Same two functions against the reference build itself (oracle/_ref) on other rows."""
    gold = np.load(os.path.join(HERE, "golden", "ref_golden.npz"))
    feats = gold["brain_small_features"]
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    pkg.api.write_features_bin(a, feats, 140.0)
    assert reference.write_bin(feats, b, 140.0) == 0
    assert open(a, "rb").read() == open(b, "rb").read()
    key = str(tmp_path / "f.key")
    pkg.api.write_features_text(key, feats, (91, 109, 91))
    assert pkg.api.read_features_text(key).tobytes() == reference.read_text(key).tobytes()
