"""Profiling driver (not a test): blur levels of the octave-0 schedule at MNI size run back to back
(a -> b -> a ..., warm L2 as in the pipeline).  Prints GPU time per level from CUDA events over the whole
chain and the per-kernel mean durations from CUPTI (torch.profiler)."""
import importlib, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
pkg = importlib.import_module("3d_sift_cuda_b200")
X, Y, Z = [int(v) for v in os.environ.get("PROF_SHAPE", "182,218,182").split(",")]
pitch = (X + 7) // 8 * 8
rng = np.random.default_rng(0)
a = torch.zeros((Z, Y, pitch), dtype=torch.float32, device="cuda")
a[:, :, :X] = torch.from_numpy(rng.random((Z, Y, X), dtype=np.float32)).cuda()
b, tmp, dog = torch.zeros_like(a), torch.zeros_like(a), torch.zeros_like(a)
sigmas = [1.5199, 1.2263, 1.5450, 1.9466, 2.4525, 3.0900]
if os.environ.get("PROF_SIGMAS"):
    sigmas = [float(v) for v in os.environ["PROF_SIGMAS"].split(",")]
e = pkg.Engine(0)
st = torch.cuda.ExternalStream(e.stream)
N = 20
out = []
for s in sigmas:
    taps = pkg.gaussian_taps(s)
    def chain(n):
        for i in range(n):
            if i % 2 == 0: e.blur3d(a, tmp, b, X, taps, dog)
            else: e.blur3d(b, tmp, a, X, taps, dog)
    chain(4); e.sync()
    with torch.cuda.stream(st):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); chain(N); e1.record(st)
    e.sync()
    us = e0.elapsed_time(e1) * 1e3 / N
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        chain(6); e.sync()
    per = collections.defaultdict(list)
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and "blur" in ev.name:
            nm = ev.name.replace("s3d::", "").replace("void ", "")
            nm = nm[:nm.index("(")] if "(" in nm else nm
            per[nm].append(ev.time_range.end - ev.time_range.start)
    ks = "  ".join("%s %.1f" % (k, sum(v) / len(v)) for k, v in per.items())
    out.append(us)
    print("%2d taps: %6.1f us/level (events, chain of %d) | kernels: %s" % (len(taps), us, N, ks))
knobs = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("S3D_"))
N0 = X * Y * Z
print("sum %.1f us | last level %.0f GB/s algorithmic (12 B/voxel) | %s" % (sum(out), 12.0 * N0 / (out[-1] * 1e-6) / 1e9, knobs))
