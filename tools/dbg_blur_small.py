import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("3d_sift_cuda_b200")
X, Y, Z = 48, 40, 36
sigma = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5199
vol = pkg.phantom.blob_phantom((X, Y, Z), 3, 20)
a = torch.from_numpy(vol).cuda()
tmp, out, dog = torch.zeros_like(a), torch.zeros_like(a), torch.zeros_like(a)
e = pkg.Engine(0)
e.blur3d(a, tmp, out, X, pkg.gaussian_taps(sigma), dog)
e.sync()
print("ok", float(out.sum()))
