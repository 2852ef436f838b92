"""Small case for compute-sanitizer --tool memcheck (not a test): one extraction per blur path on an odd-sized
volume, typed input, -2+, batch of 3; prints the row counts."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("3d_sift_cuda_b200")
vol = pkg.phantom.blob_phantom((61, 53, 47), 4, 50)
e = pkg.Engine(0)
print("rows", len(e.extract(vol)), len(e.extract(vol, pkg.Params(double_mode=1))), len(e.extract(vol, pkg.Params(descriptor=2))),
      len(e.extract_typed((vol * 30).astype(np.int16))))
e.close()
b = pkg.Batch(0, 2)
print("batch", [len(r) for r in b.extract([vol, vol[:, :, ::-1].copy(), vol])])
b.close()
