"""Profiling driver (not a test): exact kNN descriptor matching (s3d_match_device / s3d_match) of
the features of a few volumes against a database of 256 volumes' features, against the oracle's exhaustive search
(the reference's DistSqrPCs) on the host.  192 FP32 operations per descriptor pair (64 x sub, mul, add: no FMA)."""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
pkg = importlib.import_module("3d_sift_cuda_b200")
from oracle_bindings import Oracle
api = pkg.api
rng = np.random.default_rng(3)
ROWS = 1045
nA, nB, K = int(os.environ.get("PROF_NA", 8 * ROWS)), int(os.environ.get("PROF_NB", 256 * ROWS)), int(os.environ.get("PROF_K", 2))
def feats(n):
    f = np.zeros(n, api.FEATURE_DTYPE)
    f["pc"] = np.argsort(rng.random((n, 64)), axis=1).astype(np.float32)      # rank descriptors: permutations of 0..63
    return f
a, b = feats(nA), feats(nB)
e = pkg.Engine(0)
da = torch.from_numpy(a.view(np.uint8)).cuda(); db = torch.from_numpy(b.view(np.uint8)).cuda()
idx = torch.empty((nA, K), dtype=torch.int32, device="cuda"); dist = torch.empty((nA, K), dtype=torch.float32, device="cuda")
L = e.L
def run():
    st = L.s3d_match_device(e.ctx, C.c_void_p(da.data_ptr()), nA, C.c_void_p(db.data_ptr()), nB, K, C.c_void_p(idx.data_ptr()), C.c_void_p(dist.data_ptr()))
    assert st == 0, st
for _ in range(2): run()
e.sync()
st = torch.cuda.ExternalStream(e.stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 5
with torch.cuda.stream(st):
    e0.record(st)
    for _ in range(N): run()
    e1.record(st)
e.sync()
ms = e0.elapsed_time(e1) / N
pairs = float(nA) * nB
print("s3d_match_device: %d x %d descriptors, k = %d: %.2f ms = %.1f G pairs/s = %.1f T FP32 op/s (%.0f %% of the 37.2 T/s the SMs issue)" %
      (nA, nB, K, ms, pairs / ms / 1e6, pairs * 192 / ms / 1e9, 100 * pairs * 192 / ms / 1e9 / 37.2))
t0 = time.perf_counter(); hi, hd = e.match(a, b, K); th = time.perf_counter() - t0
print("s3d_match (host arrays in, host results out): %.2f ms" % (1e3 * th))
assert (hi == idx.cpu().numpy()).all()
ns = 64
O = Oracle()
t0 = time.perf_counter(); oi, od = O.knn(a["pc"][:ns], b["pc"], K); to = time.perf_counter() - t0
assert (oi == hi[:ns]).all() and (od == hd[:ns]).all()
print("oracle (one host core) on %d of the queries: %.2f s -> %.1f M pairs/s; identical neighbours and distances; GPU / core = %.0fx" %
      (ns, to, ns * nB / to / 1e6, (pairs / ms * 1e3) / (ns * nB / to)))
