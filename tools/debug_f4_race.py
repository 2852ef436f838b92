"""Debug driver: the one-kernel level against the two-kernel level on a large volume, repeated (race hunting)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("3d_sift_cuda_b200")
X, Y, Z = [int(v) for v in os.environ.get("PROF_SHAPE", "1024,1024,512").split(",")]
pitch = (X + 7) // 8 * 8
g = torch.Generator(device="cuda"); g.manual_seed(1)
a = torch.zeros((Z, Y, pitch), dtype=torch.float32, device="cuda")
a[:, :, :X] = torch.rand((Z, Y, X), generator=g, device="cuda") * 100
tmp = torch.zeros_like(a)
def run(env, sigma, n):
    for k, v in env.items(): os.environ[k] = v
    e = pkg.Engine(0)
    outs = []
    taps = pkg.gaussian_taps(sigma)
    for _ in range(n):
        b, d = torch.zeros_like(a), torch.zeros_like(a)
        torch.cuda.synchronize()      # the engine runs on its own non-blocking stream
        e.blur3d(a, tmp, b, X, taps, d); e.sync()
        outs.append((b, d))
    e.close()
    for k in env: del os.environ[k]
    return outs
for sigma in (1.2263, 1.5450, 1.9466, 2.4525):
    ref = run({"S3D_F4_MAXR": "0"}, sigma, 1)[0]
    outs = run({}, sigma, 10)
    bad = []
    for i, (b, d) in enumerate(outs):
        nb = int((b.view(torch.int32) != ref[0].view(torch.int32)).sum()); nd = int((d.view(torch.int32) != ref[1].view(torch.int32)).sum())
        bad.append((nb, nd))
        if nb:
            idx = (b.view(torch.int32) != ref[0].view(torch.int32)).nonzero()[:4].cpu().numpy()
            print("   run %d first bad (z,y,x):" % i, idx.tolist())
    print("sigma %.4f (%d taps): mismatching voxels (level, dog) per run: %s" % (sigma, len(pkg.gaussian_taps(sigma)), bad), flush=True)
