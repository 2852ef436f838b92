"""Profiling driver: is the batch throughput bound by the ONE host thread that feeds the contexts?  T host threads,
each with its own s3d_batch of C contexts, extract device-resident volumes concurrently (ctypes releases the GIL)."""
import importlib, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("3d_sift_cuda_b200")
T = int(os.environ.get("PROF_THREADS", "2")); C = int(os.environ.get("PROF_CONTEXTS", "3")); nvol = int(os.environ.get("PROF_VOLUMES", "96"))
vols = [torch.from_numpy(pkg.phantom.brain_phantom((182, 218, 182), 1 + i, 400)).cuda() for i in range(8)]
torch.cuda.synchronize()
bs = [pkg.Batch(0, C) for _ in range(T)]
prm = pkg.Params()
seq = [vols[i % 8] for i in range(nvol // T)]
for b in bs: b.extract_device(seq[:2 * C], (182, 218, 182), prm)
best = 1e9
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=lambda b=b: b.extract_device(seq, (182, 218, 182), prm)) for b in bs]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    best = min(best, (time.perf_counter() - t0) / (len(seq) * T) * 1e6)
print("host threads %d x %d contexts: %.1f us/volume = %.0f volumes/s" % (T, C, best, 1e6 / best))
