"""Profiling driver (not a test): device-resident batch throughput (s3d_batch_extract_device) at MNI size for
whatever S3D_* knobs are set -- the quantity bench.py reports as `value`."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("3d_sift_cuda_b200")
if os.environ.get("PROF_LIB"):      # A/B runs of two builds on the same box
    pkg.api.library_path = lambda: os.path.join(os.path.dirname(pkg.api.__file__), os.environ["PROF_LIB"])
nctx = int(os.environ.get("PROF_CONTEXTS", "4"))
nvol = int(os.environ.get("PROF_VOLUMES", "48"))
vols = [torch.from_numpy(pkg.phantom.brain_phantom((182, 218, 182), 1 + i, 400)).cuda() for i in range(8)]
torch.cuda.synchronize()
b = pkg.Batch(0, nctx)
prm = pkg.Params(max_octaves=int(os.environ.get("PROF_MAX_OCT", "0")))
seq = [vols[i % 8] for i in range(nvol)]
b.extract_device(seq[:2 * nctx], (182, 218, 182), prm)
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); nk, nr = b.extract_device(seq, (182, 218, 182), prm); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / nvol * 1e3)
knobs = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("S3D_") or k.startswith("PROF_"))
print("batch %d contexts: %.1f us/volume = %.0f volumes/s | rows %d | launches/volume %d | %s" % (nctx, best, 1e6 / best, nr[0], b.launches_per_volume(), knobs))
