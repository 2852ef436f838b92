"""Profiling driver (not a test): 4 launches of one blur level (x, y, z+DoG passes) at MNI size.
usage: python tools/prof_blur_one.py [sigma]   (default 3.09 -> 17 taps)"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("3d_sift_cuda_b200")
X, Y, Z = 182, 218, 182
pitch = (X + 7) // 8 * 8
sigma = float(sys.argv[1]) if len(sys.argv) > 1 else 3.09
a = torch.zeros((Z, Y, pitch), dtype=torch.float32, device="cuda"); a[:, :, :X] = torch.from_numpy(pkg.phantom.brain_phantom()).cuda()
tmp, out, dog = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
torch.cuda.synchronize()
e = pkg.Engine(0)
taps = pkg.gaussian_taps(sigma)
for _ in range(4):
    e.blur3d(a, tmp, out, X, taps, dog)
e.sync()
print("ok", len(taps))
