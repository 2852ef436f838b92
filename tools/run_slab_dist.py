"""Multi-GPU check of the z-slab mode (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_slab_dist.py [X Y Z] [double_mode]

Rank 0 compares the slab result (NCCL halo exchange per octave) with the whole-volume engine and prints
timings; exit code 1 on any difference."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

pkg = importlib.import_module("3d_sift_cuda_b200")
d = importlib.import_module("3d_sift_cuda_b200.dist")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (128, 128, 256)
dm = int(sys.argv[4]) if len(sys.argv) >= 5 else 0
CAP = 1 << 18      # keypoint capacity: a 512^3 pyramid yields far more than the default 16384
nblobs = int(os.environ.get("SLAB_BLOBS", "300"))
vol = pkg.phantom.blob_phantom(shape, 17, nblobs)
eng = pkg.Engine(local)
Z0 = shape[2] * (2 if dm == 1 else 1)
K, bounds = d.slab_plan(Z0, world)
dist.barrier()
for it in range(2):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    slab = d.extract_slab(eng, vol, rank, world, None, double_mode=dm, max_keypoints=CAP, max_features=8 * CAP)
    torch.cuda.synchronize(); dist.barrier(); t_slab = time.perf_counter() - t0
ok = True
if rank == 0:
    for it in range(2):      # second call: plan and graph already resident
        t0 = time.perf_counter()
        whole = eng.extract(vol, pkg.Params(double_mode=dm, max_keypoints=CAP, max_features=8 * CAP))
        t_whole = time.perf_counter() - t0
    ok = (len(slab) == len(whole)) and slab.tobytes() == whole.tobytes()
    print("slab mode: world %d, shape %s, double_mode %d, slab octaves K=%d, bounds %s" % (world, shape, dm, K, bounds))
    print("rows slab %d whole %d  identical %s  | slab %.1f ms (incl. host slicing + H2D), whole-volume on one GPU %.1f ms"
          % (len(slab), len(whole), ok, 1e3 * t_slab, 1e3 * t_whole))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
