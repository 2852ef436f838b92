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
