"""GPU parity tests: the CUDA path, called through the C-ABI (include/s3d.h), against the oracle
(oracle/sift3d_oracle.c) on the same seeded inputs.  Bit-exact everywhere: the kernels reproduce the
reference CPU arithmetic operation for operation, so tolerances are zero ULP for voxels, geometry and
pre-rank descriptors, and integer equality for candidate indices and ranks.  (north_star allows
<=1e-5 relative for floats; these tests hold the stricter bar.)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def to_dev(vol, pitch=None):
    import torch
    Z, Y, X = vol.shape
    pitch = pitch or ((X + 7) // 8 * 8)
    buf = np.zeros((Z, Y, pitch), np.float32)
    buf[:, :, :X] = vol
    return torch.from_numpy(buf).cuda()


def ready():
    """The engine runs on its own stream: finish torch's pending fills before handing tensors over."""
    import torch
    torch.cuda.synchronize()


def from_dev(t, X):
    return t.cpu().numpy()[:, :, :X].copy()


@pytest.mark.parametrize("sigma", [0.5, 0.95, 1.2263, 1.2490, 1.5199, 1.5450, 1.9466, 2.4525, 3.0900, 4.0])
def test_taps_bit_exact(pkg, oracle, sigma):
    assert (bits(pkg.gaussian_taps(sigma)) == bits(oracle.taps(sigma))).all()


@pytest.mark.parametrize("shape_xyz", [(48, 40, 36), (37, 29, 23), (64, 64, 64), (9, 7, 5)])
@pytest.mark.parametrize("sigma", [0.5, 1.2263, 1.5199, 1.9466, 2.4525, 3.0900, 4.0])
def test_blur_bit_exact(pkg, oracle, engine, shape_xyz, sigma):
    import torch
    vol = pkg.phantom.blob_phantom(shape_xyz, seed=3, nblobs=20)
    X = shape_xyz[0]
    want = oracle.blur(vol, sigma)
    d_in = to_dev(vol)
    d_tmp, d_out, d_dog = torch.zeros_like(d_in), torch.zeros_like(d_in), torch.zeros_like(d_in)
    ready()
    engine.blur3d(d_in, d_tmp, d_out, X, pkg.gaussian_taps(sigma), d_dog)
    engine.sync()
    assert (bits(from_dev(d_out, X)) == bits(want)).all()
    assert (bits(from_dev(d_dog, X)) == bits(oracle.dog(vol, want))).all()
    assert (bits(from_dev(d_in, X)) == bits(vol)).all(), "input must not be clobbered"
    assert float(d_out[:, :, X:].abs().sum()) == 0.0, "padding columns must stay zero"


def test_blur_unaligned_pitch_takes_scalar_path(pkg, oracle, engine):
    import torch
    vol = pkg.phantom.blob_phantom((21, 17, 13), seed=5, nblobs=8)
    d_in = torch.from_numpy(vol).cuda()   # pitch == X == 21: the reference's dense layout
    d_tmp, d_out = torch.zeros_like(d_in), torch.zeros_like(d_in)
    ready()
    engine.blur3d(d_in, d_tmp, d_out, 21, pkg.gaussian_taps(1.5199))
    engine.sync()
    assert (bits(d_out.cpu().numpy()) == bits(oracle.blur(vol, 1.5199))).all()


def test_dog_subsample_resize_bit_exact(pkg, oracle, engine):
    import torch
    vol = pkg.phantom.blob_phantom((45, 38, 33), seed=7, nblobs=25)
    vol2 = oracle.blur(vol, 1.5)
    X = 45
    a, b = to_dev(vol), to_dev(vol2)
    out = torch.zeros_like(a)
    ready()
    engine.dog(a, b, out, X)
    engine.sync()
    assert (bits(from_dev(out, X)) == bits(oracle.dog(vol, vol2))).all()
    for name, num, den in (("subsample2", 1, 2), ("halve_size", 1, 2), ("double_size", 2, 1)):
        want = getattr(oracle, {"subsample2": "subsample", "halve_size": "halve_size", "double_size": "double_size"}[name])(vol)
        oz, oy, ox = want.shape
        d_out = torch.full((oz, oy, (ox + 7) // 8 * 8), 7.0, dtype=torch.float32, device="cuda")
        ready()
        getattr(engine, name)(a, X, d_out)
        engine.sync()
        assert (bits(from_dev(d_out, ox)) == bits(want)).all(), name
        assert float(d_out[:, :, ox:].abs().sum()) == 0.0, name


def test_detect_bit_exact_and_ordered(pkg, oracle, engine):
    vol = pkg.phantom.blob_phantom((64, 56, 48), seed=11, nblobs=80)
    g1 = oracle.blur(vol, 1.5199)
    g2 = oracle.blur(g1, 1.2263)
    g3 = oracle.blur(g2, 1.5450)
    d0, d1 = oracle.dog(g1, g2), oracle.dog(g2, g3)
    want_min, want_max = oracle.detect(d0, d1)
    got_min, got_max = engine.detect(to_dev(d0), to_dev(d1), 64)
    assert len(want_min) + len(want_max) > 10
    assert got_min.tobytes() == want_min.tobytes()
    assert got_max.tobytes() == want_max.tobytes()


CASES = [
    ("blob64", lambda p: p.blob_phantom((64, 64, 64), 0, 60), 0),
    ("blob_odd", lambda p: p.blob_phantom((61, 53, 47), 4, 50), 0),
    ("blob40_double", lambda p: p.blob_phantom((40, 44, 36), 5, 30), 1),
    ("blob96_halve", lambda p: p.blob_phantom((96, 90, 100), 6, 120), -1),
    ("brain_small", lambda p: p.brain_phantom((91, 109, 91), 1, 100), 0),
]


@pytest.mark.parametrize("name,make,double_mode", CASES, ids=[c[0] for c in CASES])
def test_extract_pyramid_keypoints_features_bit_exact(pkg, oracle, engine, name, make, double_mode):
    vol = make(pkg.phantom)
    want = oracle.extract(vol, double_mode, 0, want_keypoints=True)
    feats = engine.extract(vol, pkg.Params(double_mode=double_mode, keep_patches=True))
    # pyramid levels of octave 0 and 1 against the oracle's octave builder
    g0 = engine.level(0, 0)
    og, od, _ = oracle.octave_levels(g0)
    for j in range(6):
        assert (bits(engine.level(0, j)) == bits(og[j])).all(), "G%d" % j
    for j in range(5):
        assert (bits(engine.level(0, j, dog=True)) == bits(od[j])).all(), "D%d" % j
    if engine.num_octaves() > 1:
        assert (bits(engine.level(1, 0)) == bits(oracle.subsample(og[3]))).all()
    kps = engine.keypoints()
    assert kps.tobytes() == want["keypoints"].tobytes()
    patches, prerank = engine.patches()
    assert len(feats) == len(want["features"]) and len(feats) > 0
    assert (bits(patches) == bits(want["patches"])).all()
    assert (bits(prerank) == bits(want["prerank"])).all()
    assert feats.tobytes() == want["features"].tobytes()


@pytest.mark.parametrize("descriptor", [1, 2, 3])
def test_brief_family_bit_exact(pkg, oracle, engine, descriptor):
    vol = pkg.phantom.blob_phantom((64, 64, 64), 0, 60)
    want = oracle.extract(vol, 0, descriptor)
    feats = engine.extract(vol, pkg.Params(descriptor=descriptor, keep_patches=True))
    _, prerank = engine.patches()
    assert (bits(prerank) == bits(want["prerank"])).all()
    assert feats.tobytes() == want["features"].tobytes()


def test_extract_is_deterministic_and_reusable(pkg, engine):
    vol = pkg.phantom.blob_phantom((64, 64, 64), 0, 60)
    a = engine.extract(vol)
    b = engine.extract(pkg.phantom.blob_phantom((64, 64, 64), 1, 60))
    c = engine.extract(vol)
    assert a.tobytes() == c.tobytes() and a.tobytes() != b.tobytes()


def test_errors_are_reported_not_fatal(pkg, engine):
    vol = pkg.phantom.blob_phantom((64, 64, 64), 0, 60)
    with pytest.raises(pkg.S3DError):
        engine.extract(vol, pkg.Params(descriptor=9))
    with pytest.raises(pkg.S3DError):
        engine.extract(vol, pkg.Params(max_keypoints=4, max_features=8))   # capacity exceeded, reported
    assert len(engine.extract(vol)) > 0   # the context is still usable


def test_tiny_and_empty_volumes(pkg, oracle, engine):
    flat = np.full((16, 16, 16), 3.0, np.float32)
    assert len(engine.extract(flat)) == 0 == len(oracle.extract(flat)["features"])
    tiny = pkg.phantom.blob_phantom((2, 5, 5), 1, 1)
    assert len(engine.extract(tiny)) == 0


@pytest.mark.parametrize("double_mode,nranks,shape", [(0, 2, (64, 56, 230)), (0, 3, (48, 52, 330)), (1, 2, (40, 36, 120))])
def test_slab_decomposition_bit_exact(pkg, engine, double_mode, nranks, shape):
    """z-slab mode (emulated ranks on one GPU) against the whole-volume engine: identical rows, same order."""
    import importlib
    d = importlib.import_module("3d_sift_cuda_b200.dist")
    vol = pkg.phantom.blob_phantom(shape, 17, 160)
    whole = engine.extract(vol, pkg.Params(double_mode=double_mode))
    Z0 = shape[2] * (2 if double_mode == 1 else 1)
    K, bounds = d.slab_plan(Z0, nranks)
    assert K >= 1, "test volume too thin to exercise slab mode"
    slab = d.extract_slab(engine, vol, double_mode=double_mode, emulate_ranks=nranks)
    assert len(whole) > 50
    assert len(slab) == len(whole)
    assert slab.tobytes() == whole.tobytes()


def test_octave_run_from_level0_matches_whole(pkg, engine):
    """input_is_g0 + octave_base: running octaves >= 1 from the engine's own level 0 reproduces their rows."""
    import torch
    vol = pkg.phantom.blob_phantom((96, 88, 80), 23, 120)
    whole = engine.extract(vol)
    kps = engine.keypoints()
    rk = engine.row_keypoints()
    g0 = torch.from_numpy(engine.level(1, 0)).cuda()
    Z, Y, X = g0.shape
    ready()
    engine.extract_device(g0, (X, Y, Z), pkg.Params(input_is_g0=True, octave_base=1))
    tail = engine.fetch_features()
    want = whole[kps["octave"][rk] >= 1]
    assert len(want) > 0 and tail.tobytes() == want.tobytes()


@pytest.mark.parametrize("env", [{"S3D_F4_MIN_VOXELS": "0"}, {"S3D_F4_MAXR": "8", "S3D_F4_MIN_VOXELS": "0"}, {"S3D_F4_MAXR": "0"},
                                 {"S3D_F4_MAXR": "0", "S3D_Z2_VEC": "2", "S3D_XY2_TX": "32", "S3D_XY2_TY": "48"},
                                 {"S3D_F4_MAXR": "0", "S3D_Z2_VEC": "4", "S3D_MARCH_TARGET": "200000"},
                                 {"S3D_F4_MAXR": "0", "S3D_XY2_KY": "8"}, {"S3D_F4_MAXR": "3", "S3D_F4_CTAS": "400", "S3D_F4_MIN_VOXELS": "0"}, {"S3D_F4_TY": "32", "S3D_F4_MIN_VOXELS": "0"}])
def test_every_blur_path_bit_exact(pkg, oracle, monkeypatch, env):
    """Each selectable blur path -- the one-kernel level (s3d_blur4.cuh) with 4 and 2 rows per thread and with
    short z segments, the x+y / z kernels (s3d_blur2.cuh) with 2 and 4 columns per thread, forced tiles, 8-output
    y segments and several z segments -- must give the oracle's bits for every radius of the schedule, DoG and
    zero padding included."""
    import torch
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    eng = pkg.Engine(0)
    try:
        for shape_xyz in [(48, 40, 36), (37, 29, 23), (70, 66, 41), (130, 35, 70)]:
            vol = pkg.phantom.blob_phantom(shape_xyz, seed=5, nblobs=20)
            X = shape_xyz[0]
            for sigma in (0.5, 0.95, 1.2263, 1.5199, 1.9466, 2.4525, 3.09):
                want = oracle.blur(vol, sigma)
                d_in = to_dev(vol)
                d_tmp, d_out, d_dog = torch.zeros_like(d_in), torch.zeros_like(d_in), torch.zeros_like(d_in)
                ready()
                eng.blur3d(d_in, d_tmp, d_out, X, pkg.gaussian_taps(sigma), d_dog)
                eng.sync()
                assert (bits(from_dev(d_out, X)) == bits(want)).all(), (shape_xyz, sigma)
                assert (bits(from_dev(d_dog, X)) == bits(oracle.dog(vol, want))).all(), (shape_xyz, sigma)
                assert float(d_out[:, :, X:].abs().sum()) == 0.0
        vol = pkg.phantom.blob_phantom((64, 64, 64), 0, 60)
        assert eng.extract(vol).tobytes() == oracle.extract(vol)["features"].tobytes()
    finally:
        eng.close()


@pytest.mark.parametrize("env", [{"S3D_TINY": "0"}, {"S3D_F4_WIDE_MIN_VOXELS": "0", "S3D_F4_MIN_VOXELS": "0"}, {"S3D_SERIAL": "1"},
                                 {"S3D_NO_GRAPH": "1"}, {"S3D_DETECT2_MIN_VOXELS": "0"}, {"S3D_TAIL_BLOCKS": "1,1,1"}])
def test_every_pipeline_shape_bit_exact(pkg, oracle, monkeypatch, env):
    """The launch-count and scheduling choices of the pipeline are not allowed to change a bit: the last octaves as one
    launch (tiny_octaves_kernel) or as ordinary levels, merged or per-level detection / refinement launches, the wide
    one-kernel levels inside the pipeline, serial and graph-less execution, grouped describe rows -- rows, keypoints and
    every pyramid level against the oracle on volumes whose octave chain reaches the tiny tail."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    eng = pkg.Engine(0)
    try:
        for shape, seed, nblobs in [((96, 80, 88), 7, 90), ((45, 52, 47), 3, 40)]:
            vol = pkg.phantom.blob_phantom(shape, seed, nblobs)
            want = oracle.extract(vol, 0, 0, want_keypoints=True)
            got = eng.extract(vol)
            assert len(want["features"]) > 20 and got.tobytes() == want["features"].tobytes(), (env, shape)
            assert eng.keypoints().tobytes() == want["keypoints"].tobytes(), (env, shape)
    finally:
        eng.close()


def test_asymmetric_taps_take_the_general_path(pkg, engine):
    """Product sharing needs w[j] == w[2R-j]; taps that are not symmetric must fall back and still follow
    the left-to-right sum (checked against a numpy restatement of filter_1d)."""
    import torch
    rng = np.random.default_rng(3)
    vol = rng.random((20, 24, 40), dtype=np.float32)
    taps = np.array([0.1, 0.2, 0.3, 0.25, 0.15], np.float32)
    def blur_axis(a, axis):
        a = np.moveaxis(a, axis, -1)
        pad = np.zeros(a.shape[:-1] + (a.shape[-1] + 4,), np.float32)
        pad[..., 2:-2] = a
        acc = np.zeros_like(a)
        for j in range(5):
            acc = (acc + (taps[j] * pad[..., j:j + a.shape[-1]]).astype(np.float32)).astype(np.float32)
        return np.moveaxis(acc, -1, axis)
    want = blur_axis(blur_axis(blur_axis(vol, 2), 1), 0)
    d_in = to_dev(vol)
    d_tmp, d_out = torch.zeros_like(d_in), torch.zeros_like(d_in)
    ready()
    engine.blur3d(d_in, d_tmp, d_out, 40, taps)
    engine.sync()
    assert (bits(from_dev(d_out, 40)) == bits(want)).all()


def test_batch_matches_single_volume_extraction(pkg, engine):
    """s3d_batch (several contexts in flight on one GPU): rows identical to s3d_extract per volume, input order."""
    import torch
    vols = [pkg.phantom.blob_phantom((64, 56, 48), seed, 40 + 5 * seed) for seed in range(7)]
    want = [engine.extract(v) for v in vols]
    assert sum(len(w) for w in want) > 100
    b = pkg.Batch(0, 3)
    try:
        got = b.extract(vols)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g.tobytes() == w.tobytes()
        pinned = [torch.from_numpy(v).pin_memory() for v in vols]
        got = b.extract(pinned)
        for g, w in zip(got, want):
            assert g.tobytes() == w.tobytes()
        d = [torch.from_numpy(v).cuda() for v in vols]
        torch.cuda.synchronize()
        nk, nr = b.extract_device(d, (64, 56, 48))
        assert nr == [len(w) for w in want]
        assert b.launches_per_volume() > 0
        assert b.extract([]) == []
    finally:
        b.close()


def test_batch_at_bench_shape_against_the_oracle(pkg, oracle):
    """The configuration bench.py times (BASELINE config 4): 6 contexts per GPU on MNI-sized volumes, with the launch
    shapes only batch contexts use (one x+y / one-kernel-level CTA per SM, one z segment, capped face test).  Rows of
    every volume == the oracle's, through the host-buffer entry point and the device-resident one."""
    import torch
    vols = [pkg.phantom.brain_phantom((182, 218, 182), 1 + i, 400) for i in range(7)]
    want = [oracle.extract(v)["features"] for v in vols[:3]]
    eng = pkg.Engine(0)
    try:
        want += [eng.extract(v) for v in vols[3:]]      # the single-volume engine is pinned to the oracle at this size above
    finally:
        eng.close()
    b = pkg.Batch(0, 6)
    try:
        got = b.extract(vols)
        for g, w in zip(got, want):
            assert len(w) > 500 and g.tobytes() == w.tobytes()
        d = [torch.from_numpy(v).cuda() for v in vols]
        torch.cuda.synchronize()
        nk, nr = b.extract_device(d, (182, 218, 182))
        assert nr == [len(w) for w in want]
    finally:
        b.close()


def test_plan_cache_keeps_results_identical(pkg, engine):
    """A context keeps a few plans resident (S3D_PLAN_CACHE, default 4): alternating between shapes, and
    evicting beyond the cache size, must give the same rows as the first visit of each shape."""
    shapes = [(64, 56, 48), (40, 44, 52), (64, 56, 48), (33, 47, 41), (72, 40, 36), (40, 44, 52), (50, 50, 50), (64, 56, 48)]
    first = {}
    for k, sh in enumerate(shapes):
        vol = pkg.phantom.blob_phantom(sh, 11, 40)
        rows = engine.extract(vol)
        if sh in first:
            assert rows.tobytes() == first[sh].tobytes(), (k, sh)
        else:
            first[sh] = rows
            assert len(rows) > 0


@pytest.mark.parametrize("dtype", ["uint8", "int8", "int16", "uint16", "int32", "uint32", "float64", "float32"])
def test_typed_input_matches_host_cast(pkg, engine, dtype):
    """Typed input (s3d_extract_typed): the device-side cast gives the rows of the host-side cast the reference
    does (reg_changeDatatype1, featExtract.cpp:18-77), with and without the -2+ pre-step."""
    vol = pkg.phantom.blob_phantom((56, 48, 40), 9, 40)
    if dtype in ("uint8", "int8"):
        typed = np.clip(vol * (1.0 if dtype == "uint8" else 0.5), 0, 120).astype(dtype)
    elif dtype == "float64":
        typed = vol.astype(np.float64) * 1.0000001
    elif dtype == "float32":
        typed = vol
    else:
        typed = (vol * 37.0).astype(dtype)
    as_float = typed.astype(np.float32)
    for dm in (0, 1):
        prm = pkg.Params(double_mode=dm)
        want = engine.extract(as_float, prm)
        got = engine.extract_typed(typed, prm)
        assert len(want) > 0 and got.tobytes() == want.tobytes(), (dtype, dm)


def test_batch_typed_input(pkg, engine):
    vols = [(pkg.phantom.blob_phantom((64, 56, 48), seed, 50) * 50.0).astype(np.int16) for seed in range(5)]
    want = [engine.extract(v.astype(np.float32)) for v in vols]
    b = pkg.Batch(0, 2)
    try:
        got = b.extract_typed(vols)
        for g, w in zip(got, want):
            assert len(w) > 0 and g.tobytes() == w.tobytes()
    finally:
        b.close()


@pytest.mark.parametrize("config,descriptor", [("config1_blob128", 0), ("config2_mni", 0), ("config3_mni_brief", 1),
                                               ("config3_mni_rrief", 2), ("config3_mni_nrrief", 3)])
def test_baseline_configs_full_size_bit_exact(pkg, oracle, engine, config, descriptor):
    """BASELINE.json configs 1-3 at their full sizes (128^3 blob phantom; 182x218x182 brain phantom with the
    SIFT-Rank and the three BRIEF-family descriptors): feature rows identical to the oracle, every descriptor
    a permutation of the ranks 0..63."""
    vol = pkg.phantom.blob_phantom() if config == "config1_blob128" else pkg.phantom.brain_phantom()
    want = oracle.extract(vol, 0, descriptor)["features"]
    feats = engine.extract(vol, pkg.Params(descriptor=descriptor))
    assert len(want) > 500
    assert feats.tobytes() == want.tobytes()
    assert (np.sort(feats["pc"], axis=1) == np.arange(64, dtype=np.float32)).all()


@pytest.mark.parametrize("env", [{}, {"S3D_F4_MIN_VOXELS": "0", "S3D_DETECT2_MIN_VOXELS": "0"}, {"S3D_F4_MAXR": "0"}])
def test_negative_zero_voxels_bit_exact(pkg, oracle, monkeypatch, env):
    """Masked images carry -0.0 voxels (negative value x 0).  The reference starts every tap sum from +0.0
    (GaussBlur3D.cpp:54-58), so a window of -0.0 voxels blurs to +0.0: the levels must match bit for bit."""
    import torch
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    eng = pkg.Engine(0)
    try:
        vol = pkg.phantom.blob_phantom((56, 48, 40), 12, 30) - 50.0       # blobs on a zero background
        vol[np.abs(vol) < 1.0] = 0.0
        vol[:, :24, :] *= -1.0                                            # half of the zeros become -0.0
        vol[20:30, 30:40, 10:30] = -0.0
        vol[2:12, 2:12, 30:38] = -1e-44                                  # products underflow to -0.0: still +0.0 sums
        assert np.signbit(vol[vol == 0]).any() and (~np.signbit(vol[vol == 0])).any()
        for sigma in (1.2263, 1.5199, 3.09):
            want = oracle.blur(vol, sigma)
            d_in = to_dev(vol)
            d_tmp, d_out, d_dog = torch.zeros_like(d_in), torch.zeros_like(d_in), torch.zeros_like(d_in)
            ready()
            eng.blur3d(d_in, d_tmp, d_out, 56, pkg.gaussian_taps(sigma), d_dog)
            eng.sync()
            assert (bits(from_dev(d_out, 56)) == bits(want)).all(), sigma
            assert (bits(from_dev(d_dog, 56)) == bits(oracle.dog(vol, want))).all(), sigma
        assert eng.extract(vol).tobytes() == oracle.extract(vol)["features"].tobytes()
    finally:
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 2, 3, 8, 16])
def test_match_knn_bit_exact(pkg, oracle, engine, k):
    """s3d_match: exact k nearest neighbours on the reference's DistSqrPCs -- indices and
    distances equal to the oracle's exhaustive search, ties to the lower index, short databases padded with -1/inf."""
    rng = np.random.default_rng(11)
    a = engine.extract(pkg.phantom.blob_phantom((64, 64, 64), 0, 60))
    b = engine.extract(pkg.phantom.blob_phantom((64, 64, 64), 1, 60))
    assert len(a) > 30 and len(b) > 30
    nb = min(len(b), 100)
    big = np.zeros(5000, pkg.FEATURE_DTYPE)
    big["pc"] = np.stack([rng.permutation(64) for _ in range(5000)]).astype(np.float32)
    big[100:100 + nb] = b[:nb]; big[4000:4000 + nb] = b[:nb]       # duplicated rows across database chunks: ties
    free = np.zeros(300, pkg.FEATURE_DTYPE); free["pc"] = rng.normal(size=(300, 64)).astype(np.float32)
    for fa, fb in [(a, b), (a, a), (b, big), (free, free[::-1].copy()), (a[:5], b[:3]), (a[:1], b[:1])]:
        idx, dist = engine.match(fa, fb, k)
        want_idx, want_dist = oracle.knn(fa["pc"], fb["pc"], k)
        assert idx.tobytes() == want_idx.tobytes()
        assert dist.tobytes() == want_dist.tobytes()
    idx, dist = engine.match(a, a, 1)
    assert (idx[:, 0] <= np.arange(len(a))).all() and (dist == 0).all()     # a row's nearest neighbour in its own set is itself (or an earlier twin)
    with pytest.raises(pkg.S3DError):
        engine.match(a, b, 17)


@pytest.mark.parametrize("double_mode,nslabs,shape", [(0, 2, (64, 56, 230)), (0, 3, (48, 52, 330)), (1, 2, (40, 36, 120)), (-1, 2, (80, 72, 420))])
def test_multi_slab_cabi_bit_exact(pkg, engine, double_mode, nslabs, shape):
    """s3d_multi_extract_slab (C-ABI, one host thread per slab, halos by peer copies): with every slab on the one
    GPU of the test box the whole multi-rank path runs -- slab plan, halo exchange between octaves, collapse onto
    slab 0, offset-based row merge -- and must reproduce the whole-volume rows byte for byte."""
    vol = pkg.phantom.blob_phantom(shape, 17, 160)
    whole = engine.extract(vol, pkg.Params(double_mode=double_mode))
    assert len(whole) > 50
    m = pkg.Multi([0] * nslabs)
    try:
        slab = m.extract_slab(vol, pkg.Params(double_mode=double_mode))
        assert slab.tobytes() == whole.tobytes()
        again = m.extract_slab(vol, pkg.Params(double_mode=double_mode, descriptor=2))      # resident plans are reused
        assert again.tobytes() == engine.extract(vol, pkg.Params(double_mode=double_mode, descriptor=2)).tobytes()
    finally:
        m.close()


def test_multi_slab_against_the_oracle(pkg, oracle):
    """Config 5 in small: a -2+ volume through the slab decomposition (2 slabs) against the ORACLE, not the engine."""
    vol = pkg.phantom.blob_phantom((64, 64, 128), 3, 200)
    want = oracle.extract(vol, 1, 0)["features"]
    m = pkg.Multi([0, 0])
    try:
        got = m.extract_slab(vol, pkg.Params(double_mode=1))
    finally:
        m.close()
    assert len(want) > 100
    assert got.tobytes() == want.tobytes()


def test_multi_slab_capacity_grows(pkg, engine):
    """max_keypoints = 0 sizes the capacity from the slab and a too-small explicit capacity is grown, not fatal."""
    vol = pkg.phantom.blob_phantom((64, 56, 230), 17, 160)
    whole = engine.extract(vol)
    m = pkg.Multi([0, 0])
    try:
        assert m.extract_slab(vol, pkg.Params(max_keypoints=8)).tobytes() == whole.tobytes()
    finally:
        m.close()


def test_multi_batch_matches_single(pkg, engine):
    """s3d_multi_batch_extract: volume i on GPU i mod n (here two shards on one GPU), rows in input order."""
    vols = [pkg.phantom.blob_phantom((64, 56, 48), s, 40) for s in range(7)]
    m = pkg.Multi([0, 0], contexts_per_gpu=2)
    try:
        rows = m.batch_extract(vols)
    finally:
        m.close()
    assert len(rows) == 7
    for v, r in zip(vols, rows):
        assert r.tobytes() == engine.extract(v).tobytes()


def test_large_volume_is_deterministic_and_path_independent(pkg, monkeypatch):
    """512^3 pyramid (256^3 with -2+): the one-kernel blur levels run concurrently with the other octaves' kernels.
    Repeated runs must give identical rows, equal to the x+y / z kernel path (a missing generic->async proxy fence
    before the TMA refill of a ring stage once showed up only at this scale, as sporadically wrong DoG voxels)."""
    vol = pkg.phantom.brain_phantom((256, 256, 256), 1, 1000)
    prm = pkg.Params(double_mode=1, max_keypoints=1 << 16, max_features=1 << 19)
    eng = pkg.Engine(0)
    try:
        runs = [eng.extract(vol, prm) for _ in range(3)]
    finally:
        eng.close()
    assert len(runs[0]) > 1000
    assert all(r.tobytes() == runs[0].tobytes() for r in runs)
    monkeypatch.setenv("S3D_F4_MAXR", "0")
    eng = pkg.Engine(0)
    try:
        assert eng.extract(vol, prm).tobytes() == runs[0].tobytes()
    finally:
        eng.close()


def test_bucketed_candidate_ranking_bit_exact(pkg, oracle, monkeypatch):
    """Large volumes rank their candidates inside z-plane buckets (cand_bucket_kernel) instead of over the whole
    list; forced on for small volumes here, the rows must still be the oracle's, in the oracle's order."""
    monkeypatch.setenv("S3D_BUCKET_MIN_VOXELS", "0")
    eng = pkg.Engine(0)
    try:
        for vol, dm in ((pkg.phantom.blob_phantom((64, 64, 64), 0, 60), 0), (pkg.phantom.brain_phantom((91, 109, 91), 1, 100), 0),
                        (pkg.phantom.blob_phantom((40, 44, 36), 31, 45), 1)):
            want = oracle.extract(vol, dm, 0)["features"]
            assert len(want) > 5
            assert eng.extract(vol, pkg.Params(double_mode=dm)).tobytes() == want.tobytes()
    finally:
        eng.close()
