"""Ad-hoc GPU smoke + timing (not collected by pytest)."""
import importlib, sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("3d_sift_cuda_b200")
from oracle_bindings import Oracle
O = Oracle()
e = pkg.Engine(0)
vol = pkg.phantom.blob_phantom((64, 64, 64), 0, 60)
want = O.extract(vol, 0, 0, want_keypoints=True)
feats = e.extract(vol, pkg.Params(keep_patches=True))
print("rows gpu", len(feats), "oracle", len(want["features"]), "equal", feats.tobytes() == want["features"].tobytes())
kps = e.keypoints(); print("kps", len(kps), len(want["keypoints"]), kps.tobytes() == want["keypoints"].tobytes())
print("launches", e.launch_count())
for name, v in [("blob128", pkg.phantom.blob_phantom()), ("brainB", pkg.phantom.brain_phantom())]:
    t = time.time(); w = O.extract(v); to = time.time() - t
    f = e.extract(v)
    print(name, "rows", len(f), len(w["features"]), "equal", f.tobytes() == w["features"].tobytes(), "oracle %.2fs" % to)
    d = torch.from_numpy(v).cuda()
    Z, Y, X = v.shape
    for i in range(3):
        e.extract_device(d, (X, Y, Z)); e.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(e.stream)
    n = 20
    with torch.cuda.stream(st):
        ev0.record(st)
        for i in range(n):
            e.extract_device(d, (X, Y, Z))
        ev1.record(st)
    e.sync()
    print(name, "device-resident ms/volume %.3f" % (ev0.elapsed_time(ev1) / n), "counts", e.fetch_counts())
    t = time.time()
    for i in range(n):
        f = e.extract(v)
    print(name, "host e2e ms/volume %.3f" % ((time.time() - t) / n * 1e3))
