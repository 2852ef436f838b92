"""CPU: the N>1 sharding path with the gloo backend, world_size 2 (host logic only, no kernels)."""
import importlib
import os
import sys

import numpy as np
import pytest


def _worker(rank, world, port, q):
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    d = importlib.import_module("3d_sift_cuda_b200.dist")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    vols = [np.full((2, 2, 2), float(i), np.float32) for i in range(7)]
    calls = []

    def fake_extract(v):
        calls.append(int(v[0, 0, 0]))
        return np.arange(int(v[0, 0, 0]) + 1, dtype=np.float32)   # result identifies the volume

    out = d.extract_sharded(fake_extract, vols, rank, world)
    tmax = d.max_over_ranks(1.0 + rank)
    q.put((rank, calls, None if out is None else [o.tolist() for o in out], tmax))
    dist.destroy_process_group()


def test_shard_indices(pkg):
    d = importlib.import_module("3d_sift_cuda_b200.dist")
    assert d.shard_indices(7, 0, 2) == [0, 2, 4, 6] and d.shard_indices(7, 1, 2) == [1, 3, 5]
    for world in (1, 2, 3, 8):
        allidx = sorted(i for r in range(world) for i in d.shard_indices(256, r, world))
        assert allidx == list(range(256))
    assert d.shard_indices(1, 3, 8) == []
    with pytest.raises(ValueError):
        d.shard_indices(4, 2, 2)


def test_extract_sharded_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, calls0, out0, t0), (r1, calls1, out1, t1) = res
    assert calls0 == [0, 2, 4, 6] and calls1 == [1, 3, 5]          # no volume processed twice
    assert out1 is None
    assert out0 == [list(map(float, range(i + 1))) for i in range(7)]   # input order restored on rank 0
    assert t0 == t1 == 2.0                                          # max over ranks


def test_extract_sharded_single_rank(pkg):
    d = importlib.import_module("3d_sift_cuda_b200.dist")
    out = d.extract_sharded(lambda v: v * 2, [np.ones(2), np.zeros(2)])
    assert out[0].tolist() == [2, 2] and out[1].tolist() == [0, 0]


def test_slab_plan_and_merge(pkg):
    d = importlib.import_module("3d_sift_cuda_b200.dist")
    K, b = d.slab_plan(1024, 8)            # config 5: 512^3 with -2+ -> 1024 planes over 8 GPUs
    assert K == 2 and b == [0, 128, 256, 384, 512, 640, 768, 896, 1024]
    K, b = d.slab_plan(364, 2)             # MNI with -2+
    assert K == 2 and b[0] == 0 and b[-1] == 364 and b[1] % 4 == 0
    assert all((b1 >> (K - 1)) - (b0 >> (K - 1)) >= d.SLAB_HALO for b0, b1 in zip(b[:-1], b[1:]))
    assert d.slab_plan(60, 4)[0] == 0      # too thin to split: whole-volume mode
    assert d.slab_plan(500, 1)[0] >= 1
    # merge order: octave, level, min/max, rank (= z order)
    def mk(tag, levels, maxs):
        f = np.zeros(len(levels), pkg.FEATURE_DTYPE)
        f["x"] = [tag * 100 + i for i in range(len(levels))]
        return f, np.array(levels), np.array(maxs)
    per_rank = [[mk(1, [1, 1, 2], [0, 1, 0])], [mk(2, [1, 2, 2], [0, 0, 1])]]
    merged = np.concatenate(d.merge_slab_rows(per_rank, 1))["x"].tolist()
    assert merged == [100, 200, 101, 102, 201, 202]


def _slab_gather_worker(rank, world, port, q):
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    pkg = importlib.import_module("3d_sift_cuda_b200")
    d = importlib.import_module("3d_sift_cuda_b200.dist")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)

    def mk(tag, levels, maxs):
        f = np.zeros(len(levels), pkg.FEATURE_DTYPE)
        f["x"] = [tag * 100 + i for i in range(len(levels))]
        f["pc"][:, 0] = tag
        return f, np.array(levels), np.array(maxs)

    # two slab octaves per rank, rows already in (level, min then max) order; rank 1 has an empty octave
    mine = [mk(1, [1, 1, 2, 3], [0, 1, 0, 1]), mk(3, [2, 2], [0, 1])] if rank == 0 else [mk(2, [1, 2, 2, 3, 3], [0, 0, 1, 0, 1]), mk(4, [], [])]
    per_rank = d.gather_slab_rows(mine, 2, rank, world)
    out = None
    if rank == 0:
        merged = np.concatenate(d.merge_slab_rows(per_rank, 2))
        out = (merged["x"].tolist(), merged["pc"][:, 0].tolist())
    q.put((rank, out))
    dist.destroy_process_group()


def test_slab_rows_gathered_as_bytes_gloo_world2():
    """The slab mode's row transport (counts + raw bytes, NCCL on GPUs) with gloo on the CPU: rank 0 rebuilds
    every rank's (features, level, is_max) and merges them in the reference's order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_slab_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    xs, tags = res[0]
    # octave 0: level1 min (r0, r1), level1 max (r0), level2 min (r0, r1), level2 max (r1), level3 min (r1), level3 max (r0, r1); octave 1: r0 only
    assert xs == [100, 200, 101, 102, 201, 202, 203, 103, 204, 300, 301]
    assert tags == [1, 2, 1, 1, 2, 2, 2, 1, 2, 3, 3]
    assert res[1] is None
