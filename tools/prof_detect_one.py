"""Profiling driver (not a test): the detection pass on octave-0 DoG levels of the MNI phantom."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("3d_sift_cuda_b200")
vol = pkg.phantom.brain_phantom()
Z, Y, X = vol.shape
e = pkg.Engine(0)
e.extract(vol)
pitch = (X + 7) // 8 * 8
def dev(a):
    t = torch.zeros((Z, Y, pitch), dtype=torch.float32, device="cuda"); t[:, :, :X] = torch.from_numpy(a).cuda(); return t
d1, d2 = dev(e.level(0, 1, dog=True)), dev(e.level(0, 2, dog=True))
for _ in range(4):
    mn, mx = e.detect(d1, d2, X)
print("candidates", len(mn), len(mx))
