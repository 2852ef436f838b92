"""Profiling driver (not a test): kernel breakdown of ONE large extraction (default 512^3 with -2+ = BASELINE config 5
on a single GPU) via CUPTI (torch.profiler): total time and launch count per kernel name, plus the wall clock."""
import importlib, os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
pkg = importlib.import_module("3d_sift_cuda_b200")
S = int(os.environ.get("PROF_SIZE", "512"))
dm = int(os.environ.get("PROF_DOUBLE", "1"))
blobs = int(os.environ.get("PROF_BLOBS", "8000"))
t0 = time.perf_counter()
vol = pkg.phantom.brain_phantom((S, S, S), 1, blobs) if os.environ.get("PROF_KIND", "brain") == "brain" else pkg.phantom.blob_phantom((S, S, S), 17, blobs)
print("phantom %.1f s" % (time.perf_counter() - t0), flush=True)
e = pkg.Engine(0)
prm = pkg.Params(double_mode=dm, max_keypoints=1 << 19, max_features=1 << 22)
for it in range(2):
    t0 = time.perf_counter()
    rows = e.extract(vol, prm)
    print("extract call %d: %.1f ms wall, %d rows" % (it, 1e3 * (time.perf_counter() - t0), len(rows)), flush=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    rows = e.extract(vol, prm)
    wall = 1e3 * (time.perf_counter() - t0)
evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
tot = collections.defaultdict(float); cnt = collections.Counter()
for ev in evs:
    nm = ev.name.replace("s3d::", "").replace("void ", "")
    nm = nm[:nm.index("(")] if "(" in nm else nm
    tot[nm] += (ev.time_range.end - ev.time_range.start) * 1e-3
    cnt[nm] += 1
span = (max(ev.time_range.end for ev in evs) - min(ev.time_range.start for ev in evs)) * 1e-3
print("profiled call: %.1f ms wall, device span %.1f ms, sum of kernels+copies %.1f ms" % (wall, span, sum(tot.values())))
for nm, ms in sorted(tot.items(), key=lambda kv: -kv[1])[:28]:
    print("  %9.2f ms  x%-4d %s" % (ms, cnt[nm], nm[:70]))
