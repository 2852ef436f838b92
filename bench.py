#!/usr/bin/env python
"""bench.py -- throughput of the featExtract hot path (pyramid + DoG + detection + refinement +
orientation + SIFT-Rank descriptors) on synthetic MNI-sized phantoms.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one batch of --batch (default 32) 182x218x182 fp32 volumes through the whole path on each GPU:
BASELINE.json config 4 ("batch of 256 MNI-sized volumes sharded per GPU") is exactly one step at 8 GPUs, and
per-GPU work is the same at every N (weak scaling; every rank extracts its own volumes, no data-path
collective).  The K steps of a run go through the batch entry point of the C-ABI (s3d_batch_*: --contexts
extraction contexts in flight per GPU, default 6), timed as ONE region between two CUDA events with
barrier + synchronize on both sides.  Prints ONE JSON line on rank 0:

  value      volumes/s, whole job, volumes already resident in HBM when the timed region starts
             (s3d_batch_extract_device; inputs rotate through a pool larger than L2; max over ranks)
  latency_ms_per_volume   one volume alone on the GPU, one context, L2 flushed between steps
  e2e        same metric through s3d_batch_extract with HOST buffers: pinned H2D of every volume and
             D2H of its feature rows inside the timed region
  roofline   the heaviest blur level (17 taps + fused DoG at octave-0 size): algorithmic bytes
             (read G_{j-1}, write G_j, write DoG = 12 B/voxel) / measured duration vs MEASURED_PEAKS.json;
             `levels` lists all six levels of octave 0
  cpu_baseline  the reference's own CPU path (oracle/_ref) on the host cores, bounded sample
  pcie       pinned host -> device copy bandwidth of one 28.9 MB volume, all ranks copying at the same time
             (per GPU and aggregate): the ceiling `e2e` can reach with fp32 input
  slab       (N > 1) BASELINE config 5: one 512^3 volume with -2+ (1024^3 pyramid) split into z slabs over the
             N GPUs through s3d_multi_extract_slab (one host thread per GPU inside the library, halos by peer
             copies over NVLink; called from rank 0 while the other ranks wait on a CPU barrier), ms per volume,
             rows, and whether they are identical to the single-GPU whole-volume run

--impl reference times the reference's CPU implementation (oracle/_ref, else the oracle port) with
all usable host cores on the same workload and prints the same JSON shape.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (182, 218, 182)      # X, Y, Z  (MNI)
NBLOBS = 400
WORKLOAD = "single MNI-sized 182x218x182 synthetic brain phantom, SIFT-Rank descriptor"
N0 = SHAPE[0] * SHAPE[1] * SHAPE[2]


def octave_voxels(shape):
    x, y, z = shape
    out = []
    while not (x <= 2 or y <= 2 or z <= 2):
        out.append(x * y * z)
        x, y, z = x // 2, y // 2, z // 2
    return out


def algorithmic_bytes(shape):
    """92.5 B/voxel on octave 0, 84.5 B/voxel on later octaves."""
    v = octave_voxels(shape)
    return 92.5 * v[0] + 84.5 * sum(v[1:])


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args):
    """Reference CPU arm: rank 0 only; all usable host cores, one MNI volume per worker per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_baseline as cb
    procs = cb.usable_cores()
    pool = cb.CpuPool(SHAPE, 1, NBLOBS, os.path.join(ROOT, "3d_sift_cuda_b200", "phantom.py"), procs)
    # bounded sample: one MNI volume per worker per step (~3 s); stop early once ~150 s have been spent
    budget, t_start = 150.0, time.perf_counter()
    for _ in range(min(args.warmup, 1)):
        pool.step(1)
    vols, secs, rows, timed = 0, 0.0, 0, 0
    for _ in range(args.steps):
        n, dt, rows = pool.step(1)
        vols += n
        secs += dt
        timed += 1
        if time.perf_counter() - t_start > budget:
            break
    pool.close()
    value = vols / secs
    line = {
        "impl": "reference", "metric": "volumes/sec (pyramid+detect+describe)", "value": value, "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / timed,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "gvoxels_per_s": value * N0 / 1e9,
        "config": {"workload": WORKLOAD, "shape_xyz": list(SHAPE), "descriptor": "SIFT-Rank"},
        "run": {"rows_per_volume": rows, "timed_steps": timed,
                "note": "reference CPU path (featExtract without -d): %d worker processes, one volume each per step "
                        "(a step = %d volumes: a bounded sample of the workload); %d of the %d requested steps were timed" % (procs, procs, timed, args.steps)},
        "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": procs, "kind": cb.kind(),
                         "sample": "%d steps x %d volumes (one per worker process)" % (timed, procs)},
        "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import numpy as np
    import torch
    pkg = importlib.import_module("3d_sift_cuda_b200")
    dmod = importlib.import_module("3d_sift_cuda_b200.dist")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the communicator is created; keep stdout for the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # CPU baseline first (rank 0, N=1 only), before the GPU is busy
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import cpu_baseline as cb
        procs = min(cb.usable_cores(), 16)
        pool = cb.CpuPool(SHAPE, 1, NBLOBS, os.path.join(ROOT, "3d_sift_cuda_b200", "phantom.py"), procs)
        pool.step(1)
        n, dt, _ = pool.step(1)
        n2, dt2, _ = pool.step(1)
        pool.close()
        cpu = {"value": (n + n2) / (dt + dt2), "unit": "volumes/s", "cores": procs, "kind": cb.kind(),
               "sample": "2 timed rounds x %d volumes (one MNI volume per worker process), 1 warm-up round" % procs}

    # the reference's OWN CUDA path (featExtract -d0, oracle/_ref/featExtract_ref_cuda built from the reference
    # sources for sm_100a), timed on this GPU before ours: a baseline, not a target and not an oracle
    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_cuda_baseline.py"), "--reps", "3", "--timeout", "90"],
                                 capture_output=True, text=True, timeout=400)
            r = json.loads(out.stdout.strip().splitlines()[-1])
            if "d0_best_stage_sum_s" in r:
                ref_cuda = {"value": 1.0 / r["d0_best_stage_sum_s"], "unit": "volumes/s", "kind": "reference -d0 (CUDA, sm_100a build)",
                            "stage_sum_s": r["d0_best_stage_sum_s"], "process_wall_s": r["d0_best_wall_s"],
                            "rows": r["d0_runs"][0].get("rows"), "cpu_rows": (r.get("cpu") or {}).get("rows"),
                            "single_core_cpu_stage_sum_s": (r.get("cpu") or {}).get("stage_sum_s"),
                            "sample": "best of 3 CLI runs on one MNI phantom (its run-to-run spread is large: 0.17-2.1 s); value = 1 / sum of the reference's own per-stage timers "
                                      "(pyramid, DoG, detection: excludes its CPU keypoint/descriptor stages, file I/O and CUDA start-up)"}
            else:
                ref_cuda = {"unavailable": r.get("unavailable") or str(r.get("d0_runs"))[:200]}
        except Exception as exc:      # noqa: BLE001
            ref_cuda = {"unavailable": repr(exc)[:200]}

    eng = pkg.Engine(local)
    st = torch.cuda.ExternalStream(eng.stream, device=dev)
    params = pkg.Params()
    nctx = max(1, args.contexts)
    batch = pkg.Batch(local, nctx)
    # A pool of distinct volumes per rank, larger than L2 (8 x 28.9 MB = 231 MB > 126 MB), rotated so that no
    # step can reuse a previous step's input or result; every context also owns its own 0.44 GB pyramid.
    npool = 8
    vols = [pkg.phantom.brain_phantom(SHAPE, 1 + rank * npool + i, NBLOBS) for i in range(npool)]
    X, Y, Z = SHAPE
    d_vols = [torch.from_numpy(v).to(dev) for v in vols]
    h_vols = [torch.from_numpy(v).pin_memory() for v in vols]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    torch.cuda.synchronize()

    def timed(fn):
        """device time of fn() between two events on an otherwise idle GPU, barrier + synchronize on both sides"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        return dmod.max_over_ranks(e0.elapsed_time(e1), device=dev), out

    # ---- device-resident throughput (value): K volumes through s3d_batch_extract_device -------------
    batch.extract_device([d_vols[i % npool] for i in range(max(args.warmup, 2 * nctx))], SHAPE, params)
    launches_per_volume = batch.launches_per_volume() + 1      # graph nodes + the input re-pitch kernel in front of the graph
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    nvol = args.steps * args.batch                      # volumes per GPU in every timed region
    seq = [d_vols[i % npool] for i in range(nvol)]
    total_ms, (nks, nfs) = timed(lambda: batch.extract_device(seq, SHAPE, params))
    nk, nf = nks[0], nfs[0]
    ms_per_step = total_ms / args.steps
    ms_per_volume = total_ms / nvol
    value = world * 1e3 / ms_per_volume

    # ---- latency of one volume alone on the GPU (one context, L2 flushed between steps) ---------------
    for i in range(3):
        eng.extract_device(d_vols[i % npool], SHAPE, params)
    eng.sync()
    evs = []
    with torch.cuda.stream(st):
        for i in range(30):
            flush.zero_()                                  # evict L2 (untimed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            eng.extract_device(d_vols[i % npool], SHAPE, params)
            e1.record(st)
            evs.append((e0, e1))
    eng.sync()
    lat_ms = dmod.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / len(evs), device=dev)

    # ---- end to end through the public batch call with host buffers (e2e) ----------------------------
    # s3d_batch_extract: every step's 28.9 MB volume is copied from pinned host memory (H2D) and its feature
    # rows are copied back (D2H) inside the timed region; the contexts overlap the copies with compute.
    batch.extract([h_vols[i % npool] for i in range(max(4, 2 * nctx))], params)
    hseq = [h_vols[i % npool] for i in range(nvol)]
    e2e_ms, rows = timed(lambda: batch.extract(hseq, params))
    d2h = sum(r.nbytes + 12 for r in rows)
    e2e = {"value": world * nvol / (e2e_ms * 1e-3), "unit": "volumes/s", "h2d_bytes_per_step": N0 * 4 * args.batch,
           "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps,
           "h2d_gbs_per_gpu": N0 * 4 * nvol / (e2e_ms * 1e-3) / 1e9,
           "note": "s3d_batch_extract with %d contexts per GPU: pinned H2D of every volume and D2H of its rows inside the timed region" % nctx}
    # typed input: the same phantoms stored as int16, the commonest NIfTI datatype of MNI-space
    # images -- raw voxels cross PCIe (14.4 MB per volume), the cast to float runs on the device
    i16_vols = [torch.from_numpy(np.rint(v * 128.0).astype(np.int16)).pin_memory() for v in vols]
    batch.extract_typed([i16_vols[i % npool] for i in range(max(4, 2 * nctx))], params)
    iseq = [i16_vols[i % npool] for i in range(nvol)]
    i16_ms, irows = timed(lambda: batch.extract_typed(iseq, params))
    e2e_i16 = {"value": world * nvol / (i16_ms * 1e-3), "unit": "volumes/s", "h2d_bytes_per_step": N0 * 2 * args.batch,
               "d2h_bytes_per_step": sum(r.nbytes + 12 for r in irows) // args.steps, "ms_per_step": i16_ms / args.steps,
               "h2d_gbs_per_gpu": N0 * 2 * nvol / (i16_ms * 1e-3) / 1e9,
               "rows_per_volume": len(irows[0]),
               "note": "s3d_batch_extract_typed, int16 voxels (phantom x128 rounded to integers: a different input than the "
                       "float32 runs, same shape and content), %d contexts per GPU" % nctx}
    # same thing strictly one step at a time (latency of a single featExtract-style call with host buffers)
    barrier()
    t0 = time.perf_counter()
    for i in range(30):
        eng.extract_host_async(h_vols[i % npool], params)
        eng.fetch_features()
    torch.cuda.synchronize()
    e2e["serial_ms_per_volume"] = 1e3 * dmod.max_over_ranks(time.perf_counter() - t0, device=dev) / 30

    # ---- the PCIe ceiling of e2e: pinned H2D of one volume, every rank copying at the same time ------
    dst = torch.empty_like(d_vols[0])
    for _ in range(3):
        dst.copy_(h_vols[0], non_blocking=True)
    ncopy = 40
    pcie_ms, _ = timed(lambda: [dst.copy_(h_vols[i % npool], non_blocking=True) for i in range(ncopy)])
    gbs = N0 * 4 * ncopy / (pcie_ms * 1e-3) / 1e9
    pcie = {"h2d_gbs_per_gpu": gbs, "h2d_gbs_aggregate": gbs * world, "bytes_per_copy": N0 * 4, "copies": ncopy,
            "volumes_per_s_ceiling_fp32": world * gbs * 1e9 / (N0 * 4), "volumes_per_s_ceiling_int16": world * gbs * 1e9 / (N0 * 2),
            "note": "pinned host -> device cudaMemcpyAsync of one volume, %d rank(s) copying concurrently, slowest rank" % world}
    del dst
    clocks = sampler.stop() if rank == 0 else None   # sampled across all timed regions

    # ---- roofline of the dominant stage: the blur levels of octave 0 (one-kernel level up to 13 taps; x+y kernel + z kernel at 17) ----
    # Algorithmic bytes of a level = read G_{j-1}, write G_j, write DoG = 12 B/voxel (the initial blur writes no
    # DoG: 8 B/voxel).  Each level is timed alone with CUDA events on the engine's stream, L2 flushed before
    # every launch (cold) and back to back in a -> b -> a chains (warm, as inside the pipeline).  The headline
    # entry is the heaviest level (17 taps); `levels` lists all six.  `traffic` = dram bytes read + written per
    # level from the ncu --set full capture summarised in profiles/r2_blur_level_traffic.json.
    roof = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        pitch = (X + 7) // 8 * 8
        a = torch.zeros((Z, Y, pitch), dtype=torch.float32, device=dev)
        a[:, :, :X] = d_vols[0]
        b2, tmp, dog = torch.zeros_like(a), torch.zeros_like(a), torch.zeros_like(a)
        torch.cuda.synchronize()
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_blur_level_traffic.json")))["levels"]
        except Exception:
            pass
        levels = []
        sched = [("initial", 1.5199, False), ("1", 1.2263, True), ("2", 1.5450, True), ("3", 1.9466, True),
                 ("4", 2.4525, True), ("5", 3.0900, True)]
        for name, sigma, with_dog in sched:
            taps = pkg.gaussian_taps(sigma)
            dg = dog if with_dog else None
            for _ in range(3):
                eng.blur3d(a, tmp, b2, X, taps, dg)
            eng.sync()
            reps, lev = 10, []
            with torch.cuda.stream(st):
                for _ in range(reps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    eng.blur3d(a, tmp, b2, X, taps, dg)
                    e1.record(st)
                    lev.append((e0, e1))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for k in range(reps):
                    if k % 2 == 0:
                        eng.blur3d(a, tmp, b2, X, taps, dg)
                    else:
                        eng.blur3d(b2, tmp, a, X, taps, dg)
                e1.record(st)
            eng.sync()
            cold = sum(x.elapsed_time(y) for x, y in lev) / reps
            warm = e0.elapsed_time(e1) / reps
            alg = (12.0 if with_dog else 8.0) * N0
            tr = traffic.get(name)      # ncu capture of the same stage-level call (every level run with its DoG output)
            levels.append({"level": name, "taps": len(taps), "kernels": (tr or {}).get("kernels"), "algorithmic_bytes": alg, "ms_cold": cold, "ms_warm": warm,
                           "achieved_cold": alg / (cold * 1e-3) / 1e9, "frac_cold": alg / (cold * 1e-3) / 1e9 / peak,
                           "achieved_warm": alg / (warm * 1e-3) / 1e9, "frac_warm": alg / (warm * 1e-3) / 1e9 / peak,
                           "traffic": (tr["dram_read"] + tr["dram_write"]) if tr else None})
            a[:, :, :X] = d_vols[0]
            torch.cuda.synchronize()
        top = levels[-1]
        tot_alg = sum(l["algorithmic_bytes"] for l in levels)
        tot_ms = sum(l["ms_cold"] for l in levels)
        roof = {"bound": "hbm", "achieved": top["achieved_cold"], "peak": peak, "unit": "GB/s", "frac": top["frac_cold"],
                "traffic": top["traffic"],
                "kernel": "blur level 4->5 (blur_xy2_kernel<8,16> + blur_z2_kernel<8,DoG,float2>: x, y, z passes, 17 taps, fused DoG) at 182x218x182, L2 flushed before each launch; "
                          "levels 0-4 (7-13 taps) run the one-kernel level blur_f4_kernel, see `levels`",
                "ms_per_launch": top["ms_cold"], "algorithmic_bytes": top["algorithmic_bytes"], "peak_source": peak_src,
                "levels": levels,
                "octave0_blur_chain": {"algorithmic_bytes": tot_alg, "ms": tot_ms, "achieved": tot_alg / (tot_ms * 1e-3) / 1e9,
                                       "frac": tot_alg / (tot_ms * 1e-3) / 1e9 / peak},
                "pipeline": {"algorithmic_bytes": algorithmic_bytes(SHAPE), "ms": ms_per_volume,
                             "achieved": algorithmic_bytes(SHAPE) / (ms_per_volume * 1e-3) / 1e9,
                             "frac": algorithmic_bytes(SHAPE) / (ms_per_volume * 1e-3) / 1e9 / peak,
                             "note": "whole step incl. keypoint stages vs SURVEY 8(d) algorithmic bytes"}}

    # ---- exact kNN descriptor matching of this volume's rows against a 256-volume database ------
    match = None
    if rank == 0:
        try:
            import ctypes as C
            import numpy as np
            rng = np.random.default_rng(3)
            n_a, n_b, k = nf, 256 * nf, 2
            def rank_features(n):
                f = np.zeros(n, pkg.api.FEATURE_DTYPE)
                f["pc"] = np.argsort(rng.random((n, 64)), axis=1).astype(np.float32)
                return f
            fa, fb = rank_features(n_a), rank_features(n_b)
            d_a = torch.from_numpy(fa.view(np.uint8)).to(dev); d_b = torch.from_numpy(fb.view(np.uint8)).to(dev)
            d_i = torch.empty((n_a, k), dtype=torch.int32, device=dev); d_d = torch.empty((n_a, k), dtype=torch.float32, device=dev)
            def run_match():
                st_ = eng.L.s3d_match_device(eng.ctx, C.c_void_p(d_a.data_ptr()), n_a, C.c_void_p(d_b.data_ptr()), n_b, k,
                                             C.c_void_p(d_i.data_ptr()), C.c_void_p(d_d.data_ptr()))
                assert st_ == 0, st_
            for _ in range(2): run_match()
            eng.sync()
            es = torch.cuda.ExternalStream(eng.stream)
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(es):
                m0.record(es)
                for _ in range(5): run_match()
                m1.record(es)
            eng.sync()
            mms = m0.elapsed_time(m1) / 5
            match = {"workload": "exact %d nearest neighbours of %d SIFT-Rank descriptors among %d (256 volumes), s3d_match_device" % (k, n_a, n_b),
                     "ms": mms, "pairs_per_s": n_a * n_b / (mms * 1e-3), "fp32_tops": n_a * n_b * 192 / (mms * 1e-3) / 1e12,
                     "frac_of_fp32_issue_peak": n_a * n_b * 192 / (mms * 1e-3) / 1e12 / 37.2,
                     "note": "192 separately rounded FP32 operations per descriptor pair (the reference's DistSqrPCs); peak = 148 SMs x 128 lanes x 1.965 GHz"}
            del d_a, d_b, d_i, d_d
        except Exception as exc:      # noqa: BLE001
            match = {"error": repr(exc)[:200]}

    # ---- BASELINE config 5 (N > 1): one 512^3 volume with -2+ split into z slabs over the N GPUs ------------
    slab = None
    if world > 1 and not args.no_slab:
        cpu_group = dist.new_group(backend="gloo")      # the other ranks wait on the CPU: their GPUs belong to rank 0's slabs now
        eng.close(); batch.close()
        del d_vols, h_vols, i16_vols, flush
        torch.cuda.empty_cache()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                S = args.slab_size
                big_pinned = torch.from_numpy(pkg.phantom.brain_phantom((S, S, S), 1, args.slab_blobs)).pin_memory()
                big = big_pinned.numpy()             # page-locked like the batch inputs: the slabs' H2D copies run at PCIe speed
                prm5 = pkg.Params(double_mode=1)
                m = pkg.Multi(list(range(world)))
                t0 = time.perf_counter()
                rows5 = m.extract_slab(big, prm5)          # first call: plans and graphs are built
                first_ms = 1e3 * (time.perf_counter() - t0)
                reps = []
                for _ in range(6):
                    t0 = time.perf_counter()
                    again = m.extract_slab(big, prm5)
                    reps.append(1e3 * (time.perf_counter() - t0))
                m.close()
                slab = {"workload": "one %d^3 volume with -2+ (%d^3 pyramid), %d blobs, z slabs over %d GPUs (s3d_multi_extract_slab)" % (S, 2 * S, args.slab_blobs, world),
                        "ms_per_volume": min(reps), "ms_per_volume_all": reps, "first_call_ms": first_ms, "rows": int(len(rows5)),
                        "deterministic": bool(again.tobytes() == rows5.tobytes()),
                        "timing": "host wall clock around the call: pinned host volume in, host rows out"}
                if not args.no_slab_check:
                    e1 = pkg.Engine(0)
                    p1 = pkg.Params(double_mode=1, max_keypoints=1 << 18, max_features=1 << 21)
                    e1.extract(big, p1)
                    t0 = time.perf_counter()
                    whole = e1.extract(big, p1)
                    slab["single_gpu_ms_per_volume"] = 1e3 * (time.perf_counter() - t0)
                    slab["identical_to_single_gpu"] = bool(whole.tobytes() == rows5.tobytes())
                    slab["speedup_vs_single_gpu"] = slab["single_gpu_ms_per_volume"] / slab["ms_per_volume"]
                    e1.close()
            except Exception as exc:      # noqa: BLE001
                slab = {"error": repr(exc)[:300]}
        dist.barrier(group=cpu_group)

    if rank == 0:
        line = {
            "metric": "volumes/sec (pyramid+detect+describe)", "value": value, "unit": "volumes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "gvoxels_per_s": value * N0 / 1e9,
            "config": {"workload": WORKLOAD, "shape_xyz": list(SHAPE), "descriptor": "SIFT-Rank"},
            "run": {"phantom": "brain_phantom(seed=1.., nblobs=400)",
                    "parallelism": ("volumes sharded over %d GPUs, no data-path collective; " % world if world > 1 else "single GPU; ")
                                   + "%d extraction contexts (streams) in flight per GPU (s3d_batch)" % nctx,
                    "contexts_per_gpu": nctx, "volumes_per_step_per_gpu": args.batch, "volumes_per_step": args.batch * world,
                    "step": "one batch of %d volumes per GPU (BASELINE config 4 = 256 volumes = one step at 8 GPUs)" % args.batch,
                    "l2": "inputs larger than L2: a pool of 8 distinct volumes (231 MB) is rotated and every context "
                          "owns a 0.44 GB pyramid; the one-volume latency below flushes L2 (256 MiB write) between steps",
                    "keypoints_per_volume": nk, "rows_per_volume": nf},
            "ms_per_volume": ms_per_volume, "latency_ms_per_volume": lat_ms,
            "clocks": clocks, "e2e": e2e, "e2e_int16": e2e_i16, "pcie": pcie, "gpu_launches": launches_per_volume * nvol,
            "launches_per_volume": launches_per_volume, "roofline": roof, "cpu_baseline": cpu, "ref_cuda_baseline": ref_cuda,
        }
        if match is not None:
            line["match"] = match
        if slab is not None:
            line["slab"] = slab
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    batch.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="volumes per GPU per step (config 4: 256 volumes / 8 GPUs)")
    ap.add_argument("--slab-size", type=int, default=512, help="edge of the config-5 volume (run with -2+)")
    ap.add_argument("--slab-blobs", type=int, default=8000)
    ap.add_argument("--no-slab", action="store_true", help="skip the config-5 z-slab measurement at N > 1")
    ap.add_argument("--no-slab-check", action="store_true", help="skip the single-GPU whole-volume comparison of the slab rows")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contexts", type=int, default=6, help="extraction contexts in flight per GPU")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
