"""Multi-GPU sharding for the featExtract path.

Batch mode (BASELINE.json config 4): volumes are independent, so volume i goes to rank i % world and
there is NO data-path collective; the only communication is the final gather of the (small) feature
lists to rank 0, which restores input order.  One process per GPU, ``torch.distributed`` (NCCL on
GPUs, gloo in the CPU tests) for the plumbing.
"""
import numpy as np


def shard_indices(n_items, rank, world):
    """Indices of the items rank ``rank`` owns: round-robin, like the reference running one process
    per volume would be scheduled (no volume is split)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_items, world))


def extract_sharded(extract_fn, volumes, rank=0, world=1, group=None, dst=0):
    """Run ``extract_fn(volume) -> ndarray`` on this rank's shard of ``volumes`` and gather the
    per-volume results on ``dst`` in input order (other ranks get None).

    ``volumes`` may be a sequence or a callable ``i -> volume`` with ``len`` given by ``n_items``
    attribute; only the owned indices are touched, so every rank can hold just its own data.
    """
    n = len(volumes)
    mine = shard_indices(n, rank, world)
    local = [(i, extract_fn(volumes[i])) for i in mine]
    if world == 1:
        out = [None] * n
        for i, f in local:
            out[i] = f
        return out
    import torch.distributed as dist
    gathered = [None] * world if rank == dst else None
    dist.gather_object(local, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = [None] * n
    for part in gathered:
        for i, f in part:
            out[i] = f
    assert all(o is not None for o in out)
    return out


def max_over_ranks(value, device=None, group=None):
    """Max of a python float over all ranks (timing is reported as the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


# ------------------------------------------------------------------------------------------------------
# Single large volume: z-slab decomposition (BASELINE.json config 5)
# ------------------------------------------------------------------------------------------------------
# Every rank owns a contiguous range of z planes.  Per octave it runs the engine on [halo | own | halo]
# (s3d_params.slab): candidates are only taken from owned planes, zero padding / support-box test /
# trilinear clamp use the global depth, so with a halo deeper than the dependency reach of one octave
# (blur radii 3+4+5+6+8 = 26 planes, +1 for detection/validation, + the 11^3 patch reach of about
# 3.46 * scale <= 30 planes on level 3) every owned keypoint is bit-identical to the whole-volume run.
# Between octaves each rank subsamples its own part of level 3 and refreshes the halos from its two
# neighbours (NCCL send/recv over NVLink) -- one exchange per octave, no collective in the voxel
# stages.  When slabs get thinner than the halo the remaining (small) octaves collapse onto rank 0.
SLAB_HALO = 48          # planes of level 0 kept valid around the owned range, per octave
INIT_BLUR_RADIUS = 4    # initial blur: 9 taps (7 after -2+), reference MultiScale.cpp:288-298


def slab_plan(z0_planes, world, halo=SLAB_HALO, max_octaves=12):
    """Plane ownership for ``world`` ranks of an octave-0 volume of depth ``z0_planes``.

    Returns (K, bounds): K = number of octaves run in slab mode, bounds[r] .. bounds[r+1] = planes of
    octave 0 owned by rank r (multiples of 2**K, so every 2x subsample stays inside a rank).  K is the
    largest count for which every rank still owns >= halo planes at octave K-1."""
    if world < 1:
        raise ValueError("world must be >= 1")
    best = (0, [0] + [z0_planes] * world)
    for K in range(1, max_octaves + 1):
        step = 1 << K
        bounds = [0]
        for r in range(1, world):
            bounds.append(int(round(r * z0_planes / world / step)) * step)
        bounds.append(z0_planes)
        if any(b1 <= b0 for b0, b1 in zip(bounds[:-1], bounds[1:])):
            break
        own_min = min((b1 >> (K - 1)) - (b0 >> (K - 1)) for b0, b1 in zip(bounds[:-1], bounds[1:]))
        depth = z0_planes >> (K - 1)
        if own_min < halo or depth <= 2:
            break
        best = (K, bounds)
    return best


def merge_slab_rows(per_rank, n_slab_octaves):
    """Restore the reference's output order from per-rank, per-octave results.

    per_rank[r][o] = (features, level, is_max) with one level / is_max entry per feature row, rows in the
    engine's order (level up, minima then maxima, raster).  Order of the merged list: octave, level,
    minima then maxima, then ranks in z order (= raster order, slabs are z-contiguous)."""
    # every (rank, octave) result is already grouped by (level, minima then maxima): slice it by the group
    # counts (views, no boolean masks -- at 1024^3 the rows are 216 MB) and let the caller concatenate once
    out = []
    for o in range(n_slab_octaves):
        groups = []
        for r in range(len(per_rank)):
            feats, lv, mx = per_rank[r][o]
            key = (np.asarray(lv, np.int64) - 1) * 2 + np.asarray(mx, np.int64)
            if len(key) > 1 and np.any(key[1:] < key[:-1]):
                raise ValueError("slab rows are not in (level, min/max) order")
            cnt = np.bincount(key, minlength=6)[:6] if len(key) else np.zeros(6, np.int64)
            off = np.concatenate([[0], np.cumsum(cnt)])
            groups.append((feats, off))
        for g in range(6):
            for feats, off in groups:
                if off[g + 1] > off[g]:
                    out.append(feats[off[g]:off[g + 1]])
    return out


def _run_slab_octave(engine, api, d_buf, o, z_off, z_global, own0, own1, base_params):
    """One octave of one slab on the engine; returns (features, level, is_max) per row."""
    nz, Y, X = d_buf.shape
    prm = api.Params(descriptor=base_params["descriptor"], eig_thres=base_params["eig_thres"],
                     max_keypoints=base_params.get("max_keypoints", 0), max_features=base_params.get("max_features", 0),
                     input_is_g0=(o > 0), octave_base=o, max_octaves=1,
                     slab=(z_off, z_global, own0, own1), pre_step_done=base_params["double_mode"])
    engine.extract_device(d_buf, (X, Y, nz), prm)
    feats = engine.fetch_features()
    kps = engine.keypoints()
    rk = engine.row_keypoints()
    return feats, kps["level"][rk], kps["is_max"][rk]


def _next_own_level0(engine, torch, z_off, own0, own1, dims_xyz):
    """2x2x2 mean of this rank's own part of level 3 -> its own part of the next octave's level 0."""
    X, Y, _ = dims_xyz
    n_next = (own1 >> 1) - (own0 >> 1)
    g3 = torch.empty((2 * n_next, Y, X), dtype=torch.float32, device="cuda")
    engine.copy_level_device(0, 3, own0 - z_off, own0 - z_off + 2 * n_next, g3)
    nxt = torch.empty((n_next, Y // 2, X // 2), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    engine.subsample2(g3, X, nxt)
    engine.sync()
    return nxt


def _octave0_image_slab(engine, torch, volume, double_mode, lo, hi):
    """Planes [lo, hi) of the pre-stepped (octave-0 resolution) image, as a dense device tensor."""
    Z, Y, X = volume.shape
    if double_mode == 0:
        return torch.from_numpy(np.ascontiguousarray(volume[lo:hi])).cuda()
    if double_mode == 1:     # fioDoubleSize: doubled plane 2z+dz needs original planes z, z+1 (clamped at the end)
        o0, o1 = lo // 2, min(Z, (hi - 1) // 2 + 2)
        src = torch.from_numpy(np.ascontiguousarray(volume[o0:o1])).cuda()
        dst = torch.empty((2 * (o1 - o0), 2 * Y, 2 * X), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        engine.double_size(src, X, dst)
        engine.sync()
        return dst[lo - 2 * o0: hi - 2 * o0].contiguous()
    o0, o1 = 2 * lo, 2 * hi  # fioSubSample2DCenterPixel
    src = torch.from_numpy(np.ascontiguousarray(volume[o0:o1])).cuda()
    dst = torch.empty((hi - lo, Y // 2, X // 2), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    engine.halve_size(src, X, dst)
    engine.sync()
    return dst


def _pre_stepped_dims(shape_zyx, double_mode):
    Z, Y, X = shape_zyx
    if double_mode == 1:
        return 2 * X, 2 * Y, 2 * Z
    if double_mode == -1:
        return X // 2, Y // 2, Z // 2
    return X, Y, Z


def gather_slab_rows(mine, K, rank, nranks, group=None):
    """Bring every rank's slab results to rank 0 (returns per_rank there, None elsewhere).

    ``mine[o] = (features, level, is_max)`` for the K slab octaves.  Rows travel as raw bytes in tensors (on the
    device with NCCL, on the host with gloo), not as pickled objects: per (octave, level, min/max) only the row
    COUNTS are needed to rebuild the level / is_max columns, because every rank's rows are already in
    (level, minima then maxima, raster) order."""
    import importlib
    import torch
    import torch.distributed as dist
    api = importlib.import_module("3d_sift_cuda_b200.api")
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    counts = np.zeros((K, 3, 2), np.int64)
    for o_, (f_, lv_, mx_) in enumerate(mine):
        for level in (1, 2, 3):
            for is_max in (0, 1):
                counts[o_, level - 1, is_max] = int(np.count_nonzero((lv_ == level) & (mx_ == is_max)))
    c_dev = torch.from_numpy(counts.reshape(-1)).to(dev)
    all_counts = [torch.empty_like(c_dev) for _ in range(nranks)]
    dist.all_gather(all_counts, c_dev, group=group)
    itemsize = api.FEATURE_DTYPE.itemsize
    if rank != 0:
        parts = [np.ascontiguousarray(f_).view(np.uint8).reshape(-1) for f_, _, _ in mine if len(f_)]
        if parts:
            dist.send(torch.from_numpy(np.concatenate(parts)).to(dev), 0, group)
        return None
    per_rank = [mine]
    for r in range(1, nranks):
        cr = all_counts[r].cpu().numpy().reshape(K, 3, 2)
        total = int(cr.sum())
        res_r = []
        if total:
            buf = torch.empty(total * itemsize, dtype=torch.uint8, device=dev)
            dist.recv(buf, r, group)
            flat = buf.cpu().numpy().view(api.FEATURE_DTYPE)
        else:
            flat = np.zeros(0, api.FEATURE_DTYPE)
        pos = 0
        for o_ in range(K):
            n_o = int(cr[o_].sum())
            lv_ = np.repeat(np.array([1, 1, 2, 2, 3, 3]), cr[o_].reshape(-1))
            mx_ = np.repeat(np.array([0, 1, 0, 1, 0, 1]), cr[o_].reshape(-1))
            res_r.append((flat[pos:pos + n_o], lv_, mx_))
            pos += n_o
        per_rank.append(res_r)
    return per_rank


class _Timer:
    """SLAB_TIMING=1: wall-clock per section of extract_slab (device synchronised at section ends)."""

    def __init__(self, torch):
        import os
        import time
        self.on = os.environ.get("SLAB_TIMING") == "1"
        self.torch, self.time, self.t, self.acc = torch, time, time.perf_counter(), {}

    def lap(self, name):
        if not self.on:
            return
        self.torch.cuda.synchronize()
        now = self.time.perf_counter()
        self.acc[name] = self.acc.get(name, 0.0) + (now - self.t)
        self.t = now

    def report(self, rank):
        if self.on:
            print("slab timing rank %d: " % rank + ", ".join("%s %.1f ms" % (k, 1e3 * v) for k, v in self.acc.items()), flush=True)


def extract_slab(engine, volume, rank=0, world=1, group=None, double_mode=0, descriptor=0, eig_thres=140.0,
                 halo=SLAB_HALO, emulate_ranks=None, max_keypoints=0, max_features=0):
    """featExtract of ONE volume split into z slabs over ``world`` ranks; rank 0 returns the feature rows
    in the reference's order (bit-identical to the whole-volume engine), other ranks return None.

    ``volume``: numpy (Z, Y, X) float32 visible to every rank (only the rank's slab + halo is uploaded).
    ``emulate_ranks=N`` runs N virtual ranks one after the other on this process' GPU (no
    torch.distributed): same code path except that halos are handed over in memory.
    """
    import importlib
    import torch
    api = importlib.import_module("3d_sift_cuda_b200.api")
    base = {"descriptor": descriptor, "eig_thres": eig_thres, "double_mode": double_mode,
            "max_keypoints": max_keypoints, "max_features": max_features}
    X0, Y0, Z0 = _pre_stepped_dims(volume.shape, double_mode)
    emu = emulate_ranks is not None
    nranks = emulate_ranks if emu else world
    K, bounds = slab_plan(Z0, nranks, halo)
    my_ranks = list(range(nranks)) if emu else [rank]
    if nranks == 1 or K == 0:
        if emu or rank == 0:
            return engine.extract(volume, api.Params(double_mode=double_mode, descriptor=descriptor, eig_thres=eig_thres,
                                                     max_keypoints=max_keypoints, max_features=max_features))
        return None
    if not emu:
        import torch.distributed as dist

    tm = _Timer(torch)
    own_g0 = {}      # rank -> dense device tensor of its own planes of the current octave's level 0
    results = {r: [] for r in my_ranks}
    dims = (X0, Y0, Z0)
    for o in range(K):
        Xo, Yo, Zo = dims
        own = {r: (bounds[r] >> o, (bounds[r + 1] >> o) if r + 1 < nranks else Zo) for r in range(nranks)}
        # ---- assemble [halo | own | halo] per rank
        bufs = {}
        if o == 0:
            hi_ = halo + INIT_BLUR_RADIUS
            for r in my_ranks:
                lo, hi = max(0, own[r][0] - hi_), min(Zo, own[r][1] + hi_)
                bufs[r] = (lo, _octave0_image_slab(engine, torch, volume, double_mode, lo, hi))
        else:
            if emu:
                for r in my_ranks:
                    parts, lo = [], own[r][0]
                    if r > 0:
                        parts.append(own_g0[r - 1][-halo:]); lo -= halo
                    parts.append(own_g0[r])
                    if r + 1 < nranks:
                        parts.append(own_g0[r + 1][:halo])
                    bufs[r] = (lo, torch.cat(parts).contiguous())
            else:
                r = rank
                mine = own_g0[r]
                lo_t = torch.empty((halo,) + tuple(mine.shape[1:]), dtype=torch.float32, device="cuda") if r > 0 else None
                hi_t = torch.empty((halo,) + tuple(mine.shape[1:]), dtype=torch.float32, device="cuda") if r + 1 < nranks else None
                ops = []
                if r > 0:
                    ops += [dist.P2POp(dist.isend, mine[:halo].contiguous(), r - 1, group), dist.P2POp(dist.irecv, lo_t, r - 1, group)]
                if r + 1 < nranks:
                    ops += [dist.P2POp(dist.isend, mine[-halo:].contiguous(), r + 1, group), dist.P2POp(dist.irecv, hi_t, r + 1, group)]
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
                parts, lo = [], own[r][0]
                if lo_t is not None:
                    parts.append(lo_t); lo -= halo
                parts.append(mine)
                if hi_t is not None:
                    parts.append(hi_t)
                bufs[r] = (lo, torch.cat(parts).contiguous())
        tm.lap("assemble/halo")
        # ---- run the octave, keep the own part of the next octave's level 0
        nxt = {}
        for r in my_ranks:
            z_off, d_buf = bufs[r]
            torch.cuda.synchronize()
            results[r].append(_run_slab_octave(engine, api, d_buf, o, z_off, Zo, own[r][0], own[r][1], base))
            tm.lap("octave run+fetch")
            nxt[r] = _next_own_level0(engine, torch, z_off, own[r][0], own[r][1], dims)
            tm.lap("next level0")
        own_g0 = nxt
        dims = (Xo // 2, Yo // 2, Zo // 2)

    # ---- collapse: the remaining octaves run on rank 0 from the gathered level 0 of octave K
    tail = None
    if emu:
        full = torch.cat([own_g0[r] for r in range(nranks)]).contiguous()
    else:
        full = None
        if rank == 0:
            parts = [own_g0[0]]
            for r in range(1, nranks):
                n_r = (bounds[r + 1] >> K if r + 1 < nranks else dims[2]) - (bounds[r] >> K)
                t = torch.empty((n_r, dims[1], dims[0]), dtype=torch.float32, device="cuda")
                dist.recv(t, r, group)
                parts.append(t)
            full = torch.cat(parts).contiguous()
        else:
            dist.send(own_g0[rank].contiguous(), 0, group)
    if full is not None and min(dims) > 2:
        assert tuple(full.shape) == (dims[2], dims[1], dims[0]), (full.shape, dims)
        torch.cuda.synchronize()
        engine.extract_device(full, dims, api.Params(descriptor=descriptor, eig_thres=eig_thres, input_is_g0=True,
                                                     octave_base=K, pre_step_done=double_mode,
                                                     max_keypoints=max_keypoints, max_features=max_features))
        tail = engine.fetch_features()

    tm.lap("collapse tail")
    # ---- rank 0 merges (small: feature rows only)
    if emu:
        per_rank = [results[r] for r in range(nranks)]
    else:
        per_rank = gather_slab_rows(results[rank], K, rank, nranks, group)
        if rank != 0:
            tm.lap("send rows")
            tm.report(rank)
            return None
        tm.lap("recv rows")
    rows = merge_slab_rows(per_rank, K)
    if tail is not None and len(tail):
        rows.append(tail)
    if not rows:
        return np.zeros(0, api.FEATURE_DTYPE)
    out = np.concatenate(rows)
    tm.lap("merge")
    tm.report(rank)
    return out
