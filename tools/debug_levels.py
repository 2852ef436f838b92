"""Debug driver: which pyramid level differs between repeated extractions of one large volume (race hunting)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("3d_sift_cuda_b200")
S = int(os.environ.get("PROF_SIZE", "512"))
vol = pkg.phantom.brain_phantom((S, S, S), 1, int(os.environ.get("PROF_BLOBS", "8000")))
prm = pkg.Params(double_mode=1, max_keypoints=1 << 18, max_features=1 << 21)
e = pkg.Engine(0)
octs = [int(v) for v in os.environ.get("DBG_OCTS", "1,2").split(",")]
def sums():
    out = {}
    for o in octs:
        dim = 2 * S >> o
        buf = torch.empty((dim, dim, dim), dtype=torch.float32, device="cuda")
        for dog in (0, 1):
            for lv in range(5 if dog else 6):
                torch.cuda.synchronize()
                e.copy_level_device(o, lv, 0, dim, buf, dog=bool(dog)); e.sync()
                out[(o, dog, lv)] = buf.view(torch.int32).to(torch.int64).sum(dim=(1, 2)).cpu().numpy()
    return out
ref = None
for run in range(int(os.environ.get("DBG_RUNS", "5"))):
    rows = e.extract(vol, prm)
    s = sums()
    if ref is None:
        ref = s; print("run 0: %d rows" % len(rows), flush=True); continue
    bad = [(k, np.nonzero(s[k] != ref[k])[0]) for k in sorted(s) if (s[k] != ref[k]).any()]
    print("run %d: %d rows; differing levels: %s" % (run, len(rows), [(k, z[:6].tolist(), len(z)) for k, z in bad][:8]), flush=True)
