"""Debug driver: repeated extractions of one large volume must give identical rows; compares blur paths too."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("3d_sift_cuda_b200")
S = int(os.environ.get("PROF_SIZE", "256"))
dm = int(os.environ.get("PROF_DOUBLE", "1"))
vol = pkg.phantom.brain_phantom((S, S, S), 1, int(os.environ.get("PROF_BLOBS", "1000")))
prm = pkg.Params(double_mode=dm, max_keypoints=1 << 18, max_features=1 << 21)
res = {}
for cfg in os.environ.get("DBG_CFGS", ";S3D_F4_MAXR=0;S3D_DETECT2_MIN_VOXELS=0").split(";"):
    for kv in [c for c in cfg.split(",") if c]:
        k, v = kv.split("="); os.environ[k] = v
    e = pkg.Engine(0)
    runs = [e.extract(vol, prm) for _ in range(int(os.environ.get("DBG_RUNS", "4")))]
    kps = e.keypoints()
    e.close()
    for kv in [c for c in cfg.split(",") if c]:
        del os.environ[kv.split("=")[0]]
    print("[%s] rows per run %s, identical to run 0: %s" % (cfg or "default", [len(r) for r in runs], [r.tobytes() == runs[0].tobytes() for r in runs]), flush=True)
    res[cfg] = runs[0]
    if any(len(r) != len(runs[0]) for r in runs):
        a, b = runs[0], [r for r in runs if len(r) != len(runs[0])][0]
        sa = set(map(bytes, a.view(np.uint8).reshape(len(a), -1))); sb = set(map(bytes, b.view(np.uint8).reshape(len(b), -1)))
        only = [np.frombuffer(x, pkg.FEATURE_DTYPE)[0] for x in list(sa ^ sb)[:6]]
        for o in only: print("   differing row: x %.2f y %.2f z %.2f scale %.2f flag %x" % (o["x"], o["y"], o["z"], o["scale"], o["flag"]))
print({k or "default": len(v) for k, v in res.items()})
