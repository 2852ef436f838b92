"""Profiling driver (not a test): fused blur level timing vs CTA target (S3D_FUSED_CTAS), MNI size."""
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
N0 = X * Y * Z
for spec in (sys.argv[1:] or ["1:0"]):
    fused, ctas = spec.split(":")
    os.environ["S3D_FUSED"] = fused; os.environ["S3D_FUSED_CTAS"] = ctas
    e = pkg.Engine(0)
    st = torch.cuda.ExternalStream(e.stream)
    row = []
    for s in sigmas:
        taps = pkg.gaussian_taps(s)
        for _ in range(3):
            e.blur3d(a, tmp, out, X, taps, dog)
        e.sync()
        for cold in (True, False):
            ms = 0.0; reps = 10
            with torch.cuda.stream(st):
                for _ in range(reps):
                    if cold: flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st); e.blur3d(a, tmp, out, X, taps, dog); e1.record(st)
                    e.sync(); ms += e0.elapsed_time(e1)
            row.append(ms / reps * 1e3)
    cold, warm = row[0::2], row[1::2]
    print("fused=%s ctas=%4s | cold " % (fused, ctas) + " ".join("%5.1f" % t for t in cold) + " sum %6.1f | warm " % sum(cold) + " ".join("%5.1f" % t for t in warm) + " sum %6.1f us" % sum(warm),
          "| 17-tap cold: %.0f GB/s alg" % (12.0 * N0 / (cold[-1] * 1e-6) / 1e9))
    e.close()
