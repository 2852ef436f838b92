"""Profiling driver (not a test): where does the time of one extraction go outside the kernels?  Compares
per-step CUDA-event time with the CPU running ahead (bench.py style) and with a sync per step, with and
without graph replay, and times the bare cudaGraphLaunch on the CPU."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("3d_sift_cuda_b200")
vol = pkg.phantom.brain_phantom()
Z, Y, X = vol.shape
e = pkg.Engine(0)
d = torch.from_numpy(vol).cuda(); torch.cuda.synchronize()
st = torch.cuda.ExternalStream(e.stream)
for _ in range(5):
    e.extract_device(d, (X, Y, Z)); e.sync()
N = 30
# (a) run ahead
evs = []
t0 = time.perf_counter()
with torch.cuda.stream(st):
    for _ in range(N):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); e.extract_device(d, (X, Y, Z)); e1.record(st); evs.append((e0, e1))
t_cpu = (time.perf_counter() - t0) / N * 1e6
e.sync()
t_wall = (time.perf_counter() - t0) / N * 1e6
a = sorted(x.elapsed_time(y) * 1e3 for x, y in evs)
# (b) sync per step
b = []
with torch.cuda.stream(st):
    for _ in range(N):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); e.extract_device(d, (X, Y, Z)); e1.record(st); e.sync(); b.append(e0.elapsed_time(e1) * 1e3)
b.sort()
print("run-ahead: median %.1f us/step (events), CPU enqueue %.1f us/step, wall %.1f us/step | sync per step: median %.1f us" % (a[N // 2], t_cpu, t_wall, b[N // 2]))

if os.environ.get("S3D_STAMPS") == "1":
    import ctypes, numpy as np
    L = e.L
    L.s3d_debug_stamps.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    out = np.zeros(4, dtype=np.uint64)
    rows = []
    for _ in range(10):
        e.extract_device(d, (X, Y, Z)); e.sync()
        L.s3d_debug_stamps(e.ctx, out.ctypes.data)
        t = out.astype(np.int64)
        rows.append((t[1] - t[0], t[2] - t[1], t[3] - t[2], t[3] - t[0]))
    rows = np.array(rows) / 1e3
    print("stamps (us, median of 10): pad %.1f | pad end -> first graph node %.1f | graph %.1f | total %.1f" % tuple(np.median(rows, axis=0)))
