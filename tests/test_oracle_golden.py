"""CPU: the oracle (oracle/sift3d_oracle.c) against golden vectors minted from the reference's own
code (tests/golden/make_golden.py).  Bit-exact."""
import hashlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_golden.npz"))

CASES = {
    "blob64": ("blob", (64, 64, 64), 0, 60, 0),
    "blob_odd": ("blob", (61, 53, 47), 4, 50, 0),
    "blob40_double": ("blob", (40, 44, 36), 5, 30, 1),
    "blob96_halve": ("blob", (96, 90, 100), 6, 120, -1),
    "brain_small": ("brain", (91, 109, 91), 1, 100, 0),
}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.float32).tobytes()).hexdigest()


def make_volume(pkg, case):
    kind, shape, seed, nblobs, _ = CASES[case]
    fn = pkg.phantom.blob_phantom if kind == "blob" else pkg.phantom.brain_phantom
    return fn(shape, seed, nblobs)


@pytest.mark.parametrize("sigma", [0.5, 0.95, 1.2263, 1.2490, 1.5199, 1.5450, 1.9466, 2.4525, 3.0900, 4.0])
def test_taps(oracle, sigma):
    want = GOLD["taps_%g" % sigma]
    got = oracle.taps(sigma)
    assert got.tobytes() == want.tobytes()


def test_schedule_tap_counts(oracle):
    # 9 | 7, 9, 11, 13, 17 taps; 7 for the -2+ initial blur
    assert [len(oracle.taps(s)) for s in (1.5199, 1.2263, 1.5450, 1.9466, 2.4525, 3.0900, 1.2490)] == [9, 7, 9, 11, 13, 17, 7]


def test_voxel_stages(pkg, oracle):
    vol = pkg.phantom.blob_phantom((48, 40, 36), 3, 40)
    assert digest(vol) == str(GOLD["voxel_input_sha"]), "phantom generator changed"
    g = [oracle.blur(vol, 1.5199)]
    for s in (1.2263, 1.5450):
        g.append(oracle.blur(g[-1], s))
    d0, d1 = oracle.dog(g[0], g[1]), oracle.dog(g[1], g[2])
    got = {"blur0": g[0], "blur1": g[1], "blur2": g[2], "dog0": d0, "dog1": d1, "subsample": oracle.subsample(g[2]),
           "double": oracle.double_size(vol), "halve": oracle.halve_size(vol)}
    for name, arr in got.items():
        assert digest(arr) == str(GOLD["voxel_%s_sha" % name]), name
        assert arr[arr.shape[0] // 2].tobytes() == GOLD["voxel_%s_plane" % name].tobytes(), name
    mn, mx = oracle.detect(d0, d1)
    assert mn.tobytes() == GOLD["voxel_detect_min"].tobytes()
    assert mx.tobytes() == GOLD["voxel_detect_max"].tobytes()
    # raster order is part of the contract
    key = lambda c: (c["z"].astype(np.int64) * 40 + c["y"]) * 48 + c["x"]
    assert (np.diff(key(mn)) > 0).all() and (np.diff(key(mx)) > 0).all()


@pytest.mark.parametrize("case", list(CASES))
def test_extract(pkg, oracle, case):
    r = oracle.extract(make_volume(pkg, case), CASES[case][4], 0)
    assert r["features"].tobytes() == GOLD["%s_features" % case].tobytes()
    assert r["prerank"].tobytes() == GOLD["%s_prerank" % case].tobytes()
    assert digest(r["patches"]) == str(GOLD["%s_patches_sha" % case])
    # descriptors are permutations of 0..63 (rank transform)
    assert (np.sort(r["features"]["pc"], axis=1) == np.arange(64, dtype=np.float32)).all()
    assert set(np.unique(r["features"]["flag"])) <= {0, 16, 32, 48}


@pytest.mark.parametrize("desc,name", [(1, "brief"), (2, "rrief"), (3, "nrrief")])
def test_brief_family(pkg, oracle, desc, name):
    r = oracle.extract(make_volume(pkg, "blob64"), 0, desc)
    assert r["features"].tobytes() == GOLD["blob64_%s_features" % name].tobytes()
    assert r["prerank"].tobytes() == GOLD["blob64_%s_prerank" % name].tobytes()
    if desc == 1:
        assert set(np.unique(r["prerank"])) <= {0.0, 1.0}


def test_empty_and_degenerate(oracle):
    assert len(oracle.extract(np.full((16, 16, 16), 3.0, np.float32))["features"]) == 0
    assert len(oracle.extract(np.zeros((5, 5, 2), np.float32))["features"]) == 0
