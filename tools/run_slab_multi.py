"""Config 5 through the C-ABI (s3d_multi_extract_slab) from one process: python tools/run_slab_multi.py [n_gpus] [size] [blobs] [phantom]
Set S3D_SLAB_TIMING=1 for the per-slab phase breakdown on stderr."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("3d_sift_cuda_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
blobs = int(sys.argv[3]) if len(sys.argv) > 3 else 8000
kind = sys.argv[4] if len(sys.argv) > 4 else "brain"
vol = pkg.phantom.brain_phantom((S, S, S), 1, blobs) if kind == "brain" else pkg.phantom.blob_phantom((S, S, S), 17, blobs)
if os.environ.get("SLAB_PIN", "1") == "1":      # page-locked input: the slabs' H2D copies run at PCIe speed
    import torch
    pinned = torch.from_numpy(vol).pin_memory()
    vol = pinned.numpy()
m = pkg.Multi(list(range(n)))
prm = pkg.Params(double_mode=1)
for it in range(4):
    t0 = time.perf_counter()
    rows = m.extract_slab(vol, prm)
    print("call %d: %.1f ms, %d rows" % (it, 1e3 * (time.perf_counter() - t0), len(rows)), flush=True)
m.close()
if os.environ.get("SLAB_CHECK", "1") == "1":
    e = pkg.Engine(0)
    p1 = pkg.Params(double_mode=1, max_keypoints=1 << 19, max_features=1 << 22)
    for it in range(2):
        t0 = time.perf_counter(); whole = e.extract(vol, p1); t1 = time.perf_counter() - t0
    print("single GPU: %.1f ms, %d rows, identical %s" % (1e3 * t1, len(whole), whole.tobytes() == rows.tobytes()))
