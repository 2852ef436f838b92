"""CPU: the oracle against the reference's own sources (oracle/_ref), live, on seeded inputs.
Skipped when oracle/_ref was not built (no /root/reference at build time)."""
import numpy as np
import pytest


def same(a, b):
    return a.shape == b.shape and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("sigma", [0.3, 0.5, 0.95, 1.0, 1.2263, 1.5199, 2.0, 3.09, 5.0])
def test_taps(oracle, reference, sigma):
    assert same(oracle.taps(sigma), reference.taps(sigma))


@pytest.mark.parametrize("shape", [(31, 17, 9), (40, 36, 48)])
def test_voxel_stages(pkg, oracle, reference, shape):
    vol = pkg.phantom.blob_phantom(shape, 9, 15)
    for s in (0.5, 1.5199, 3.09):
        assert same(oracle.blur(vol, s), reference.blur(vol, s))
    b = oracle.blur(vol, 1.3)
    assert same(oracle.dog(vol, b), reference.dog(vol, b))
    assert same(oracle.subsample(vol), reference.subsample(vol))
    assert same(oracle.double_size(vol), reference.double_size(vol))
    assert same(oracle.halve_size(vol), reference.halve_size(vol))
    c = oracle.blur(b, 1.6)
    om, ox = oracle.detect(oracle.dog(vol, b), oracle.dog(b, c))
    rm, rx = reference.detect(oracle.dog(vol, b), oracle.dog(b, c))
    assert same(om, rm) and same(ox, rx)


@pytest.mark.parametrize("double_mode,desc", [(0, 0), (1, 0), (-1, 0), (0, 1), (0, 2), (0, 3)])
def test_extract(pkg, oracle, reference, double_mode, desc):
    shape = (72, 66, 60) if double_mode < 0 else (44, 40, 36) if double_mode > 0 else (56, 60, 52)
    vol = pkg.phantom.blob_phantom(shape, 21 + desc, 45)
    a, b = oracle.extract(vol, double_mode, desc), reference.extract(vol, double_mode, desc)
    assert len(b["features"]) > 0
    for k in ("features", "patches", "prerank"):
        assert same(a[k], b[k]), k


@pytest.mark.parametrize("k", [1, 2, 5])
def test_knn_on_the_reference_distance(pkg, oracle, reference, k):
    """Exhaustive kNN: the restatement of Feature3DInfo::DistSqrPCs + (distance, index) order against the same
    search driven by the reference's own member function; rank descriptors (many ties) and free floats."""
    rng = np.random.default_rng(5)
    ranks = lambda n: np.stack([rng.permutation(64) for _ in range(n)]).astype(np.float32)
    for pa, pb in [(ranks(40), ranks(90)), (rng.normal(size=(33, 64)).astype(np.float32), rng.normal(size=(70, 64)).astype(np.float32)),
                   (ranks(7), ranks(3))]:
        pb[1] = pb[0]                                      # exact duplicates in the database: ties go to the lower index
        fa = np.zeros(len(pa), pkg.FEATURE_DTYPE); fa["pc"] = pa
        fb = np.zeros(len(pb), pkg.FEATURE_DTYPE); fb["pc"] = pb
        oi, od = oracle.knn(pa, pb, k)
        ri, rd = reference.knn(fa, fb, k)
        assert same(oi, ri) and same(od, rd)
