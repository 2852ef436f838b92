"""Profiling driver (not a test): per-phase cycle split of the keypoint kernels, using the
-DS3D_PHASE_TIMERS build (make -C 3d_sift_cuda_b200/csrc ../lib3dsift_b200_prof.so)."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("3d_sift_cuda_b200")
api = pkg.api
api.library_path = lambda: os.path.join(os.path.dirname(api.__file__), "lib3dsift_b200_prof.so")
L = api.load_library()
vol = pkg.phantom.brain_phantom()
Z, Y, X = vol.shape
e = pkg.Engine(0)
d = torch.from_numpy(vol).cuda(); torch.cuda.synchronize()
buf = (C.c_ulonglong * 32)()
for it in range(3):
    e.extract_device(d, (X, Y, Z)); e.sync()
    L.s3d_debug_phase_cycles(buf)
nk, nf = e.fetch_counts()
names = {0: "o:load kp", 1: "o:gather", 2: "o:normalize", 3: "o:grad+tensor", 4: "o:svd", 5: "o:contrib1", 6: "o:splat1", 7: "o:blur1",
         8: "o:peaks1", 9: "o:contrib2", 10: "o:splat2", 11: "o:blur2", 12: "o:peaks2", 13: "o:rots", 22: "s:rank (match.any)", 23: "s:scan", 24: "s:scatter", 25: "s:bin sums", 16: "d:setup", 17: "d:gather/load",
         18: "d:normalize", 19: "d:grad+bins", 20: "d:accumulate", 21: "d:norm+rank+write"}
tot_o = sum(buf[i] for i in range(16)); tot_d = sum(buf[i] for i in range(16, 22))
print("keypoints", nk, "rows", nf, "orient cycles/kp %.0f" % (tot_o / max(nk, 1)), "describe cycles/row %.0f" % (tot_d / max(nf, 1)))
for i in range(32):
    if buf[i]:
        tot = tot_d if 16 <= i < 22 else tot_o
        per = buf[i] / (nf if 16 <= i < 22 else nk)
        print("%-22s %12d cycles  %5.1f%%  %8.0f cycles per %s" % (names.get(i, str(i)), buf[i], 100.0 * buf[i] / tot, per, "row" if 16 <= i < 22 else "kp"))
