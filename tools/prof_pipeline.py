"""Profiling driver (not a test): device-resident ms/volume of the whole graph-replayed extraction at MNI
size, for whatever S3D_* environment knobs are set (A/B runs of kernel variants)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("3d_sift_cuda_b200")
if os.environ.get("PROF_LIB"):      # A/B runs of two builds on the same box
    pkg.api.library_path = lambda: os.path.join(os.path.dirname(pkg.api.__file__), os.environ["PROF_LIB"])
vol = pkg.phantom.brain_phantom()
Z, Y, X = vol.shape
e = pkg.Engine(0)
d = torch.from_numpy(vol).cuda(); torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.ExternalStream(e.stream)
for _ in range(5):
    e.extract_device(d, (X, Y, Z)); e.sync()
ts = []
with torch.cuda.stream(st):
    for _ in range(int(os.environ.get("PROF_REPS", "30"))):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); e.extract_device(d, (X, Y, Z)); e1.record(st)
        e.sync()
        ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
knobs = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("S3D_"))
print("pipeline us/volume: median %.1f  min %.1f  max %.1f | counts %s | %s" % (ts[len(ts) // 2], ts[0], ts[-1], e.fetch_counts(), knobs))
