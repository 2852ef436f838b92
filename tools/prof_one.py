"""Profiling driver (not a test): N device-resident extractions of the MNI phantom.
usage: python tools/prof_one.py [n_extractions] [blob128|brainB]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("3d_sift_cuda_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
which = sys.argv[2] if len(sys.argv) > 2 else "brainB"
vol = pkg.phantom.brain_phantom() if which == "brainB" else pkg.phantom.blob_phantom()
Z, Y, X = vol.shape
e = pkg.Engine(0)
d = torch.from_numpy(vol).cuda()
torch.cuda.synchronize()
for i in range(n):
    e.extract_device(d, (X, Y, Z))
    e.sync()
print("counts", e.fetch_counts(), "launches", e.launch_count())
