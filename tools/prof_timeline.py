"""Profiling driver (not a test): kernel timeline of one graph-replayed extraction via CUPTI (torch.profiler).
Prints every kernel of the last extraction with its start offset, duration and stream."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
pkg = importlib.import_module("3d_sift_cuda_b200")
if os.environ.get("PROF_LIB"):      # A/B runs of two builds on the same box
    pkg.api.library_path = lambda: os.path.join(os.path.dirname(pkg.api.__file__), os.environ["PROF_LIB"])
vol = pkg.phantom.brain_phantom()
Z, Y, X = vol.shape
e = pkg.Engine(0)
d = torch.from_numpy(vol).cuda(); torch.cuda.synchronize()
for _ in range(5):
    e.extract_device(d, (X, Y, Z)); e.sync()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        e.extract_device(d, (X, Y, Z)); e.sync()
evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda ev: ev.time_range.start)
# split into runs at pad_rows
starts = [i for i, ev in enumerate(evs) if "pad_rows" in ev.name]
last = evs[starts[-1]:]
t0 = last[0].time_range.start
end = max(ev.time_range.end for ev in last)
print("kernels %d, span %.1f us" % (len(last), end - t0))
for ev in last:
    nm = ev.name.replace("s3d::", "").replace("void ", "")
    nm = nm[:nm.index("(")] if "(" in nm else nm
    print("%8.1f +%7.1f  %s" % (ev.time_range.start - t0, ev.time_range.end - ev.time_range.start, nm[:60]))
