"""Profiling driver (not a test): time every blur level of the octave schedule at MNI size through the
stage-level C-ABI call, for several march-pass segmentations (S3D_MARCH_TARGET threads in flight)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("3d_sift_cuda_b200")
X, Y, Z = 182, 218, 182
pitch = (X + 7) // 8 * 8
vol = pkg.phantom.brain_phantom()
a = torch.zeros((Z, Y, pitch), dtype=torch.float32, device="cuda"); a[:, :, :X] = torch.from_numpy(vol).cuda()
tmp, out, dog = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sigmas = [1.5199, 1.2263, 1.5450, 1.9466, 2.4525, 3.0900]
if os.environ.get("PROF_SIGMAS"):
    sigmas = [float(v) for v in os.environ["PROF_SIGMAS"].split(",")]
REPS = int(os.environ.get("PROF_REPS", "10"))
WARM = int(os.environ.get("PROF_WARM", "3"))
targets = [int(t) for t in sys.argv[1:]] or [0]
N0 = X * Y * Z
for tgt in targets:
    os.environ["S3D_MARCH_TARGET"] = str(tgt)
    e = pkg.Engine(0)
    st = torch.cuda.ExternalStream(e.stream)
    row = []
    for s in sigmas:
        taps = pkg.gaussian_taps(s)
        for _ in range(WARM):
            e.blur3d(a, tmp, out, X, taps, dog)
        e.sync()
        ms = 0.0
        reps = REPS
        with torch.cuda.stream(st):
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st); e.blur3d(a, tmp, out, X, taps, dog); e1.record(st)
                e.sync()
                ms += e0.elapsed_time(e1)
        row.append(ms / reps * 1e3)
    print("target %8d | " % tgt + " ".join("%2d taps %6.1f us" % (len(pkg.gaussian_taps(s)), t) for s, t in zip(sigmas, row)) + " | sum %.1f us" % sum(row),
          "| 17-tap level: %.0f GB/s algorithmic" % (12.0 * N0 / (row[-1] * 1e-6) / 1e9))
    e.close()
