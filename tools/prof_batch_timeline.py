"""Profiling driver (not a test): kernel time per volume in batch mode (s3d_batch, N contexts in flight) from
CUPTI.  Sums each kernel's durations over a batch and divides by the number of volumes: with overlap the sum
exceeds the wall time; the table shows which kernels own the GPU time when the GPU is kept full."""
import importlib, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
pkg = importlib.import_module("3d_sift_cuda_b200")
nctx = int(os.environ.get("PROF_CONTEXTS", "4"))
nvol = int(os.environ.get("PROF_VOLUMES", "16"))
vols = [torch.from_numpy(pkg.phantom.brain_phantom((182, 218, 182), 1 + i, 400)).cuda() for i in range(8)]
torch.cuda.synchronize()
b = pkg.Batch(0, nctx)
seq = [vols[i % 8] for i in range(nvol)]
b.extract_device(seq, (182, 218, 182))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    b.extract_device(seq, (182, 218, 182))
    torch.cuda.synchronize()
evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(ev.time_range.start for ev in evs); t1 = max(ev.time_range.end for ev in evs)
per = collections.defaultdict(lambda: [0.0, 0])
for ev in evs:
    nm = ev.name.replace("s3d::", "").replace("void ", "")
    nm = nm[:nm.index("(")] if "(" in nm else nm
    per[nm][0] += ev.time_range.end - ev.time_range.start
    per[nm][1] += 1
print("batch of %d volumes, %d contexts: wall %.1f us/volume (under CUPTI), kernel-time sum %.1f us/volume" %
      (nvol, nctx, (t1 - t0) / nvol, sum(v[0] for v in per.values()) / nvol))
for nm, (us, n) in sorted(per.items(), key=lambda kv: -kv[1][0]):
    print("%9.1f us/volume  %5.1f launches/volume  %7.1f us avg  %s" % (us / nvol, n / nvol, us / n, nm))
