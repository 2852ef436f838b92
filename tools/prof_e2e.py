"""Profiling driver (not a test): host-buffer batch throughput (s3d_batch_extract / _typed) at MNI size for
whatever S3D_* knobs are set -- the quantity bench.py reports as `e2e`."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("3d_sift_cuda_b200")
nctx = int(os.environ.get("PROF_CONTEXTS", "4"))
nvol = int(os.environ.get("PROF_VOLUMES", "64"))
typed = os.environ.get("PROF_INT16") == "1"
vols = [pkg.phantom.brain_phantom((182, 218, 182), 1 + i, 400) for i in range(8)]
if typed:
    h = [torch.from_numpy(np.rint(v * 128.0).astype(np.int16)).pin_memory() for v in vols]
else:
    h = [torch.from_numpy(v).pin_memory() for v in vols]
b = pkg.Batch(0, nctx)
run = b.extract_typed if typed else b.extract
seq = [h[i % 8] for i in range(nvol)]
run(seq[:2 * nctx])
best = 1e9
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    rows = run(seq)
    torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t) / nvol * 1e6)
knobs = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("S3D_") or k.startswith("PROF_"))
print("e2e %s, %d contexts: %.1f us/volume = %.0f volumes/s, %.1f GB/s H2D | rows %d | %s" %
      ("int16" if typed else "fp32", nctx, best, 1e6 / best, h[0].numel() * h[0].element_size() / best / 1e3, len(rows[0]), knobs))
