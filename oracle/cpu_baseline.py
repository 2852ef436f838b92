"""Timed CPU baseline of the featExtract path (TEST / BENCH INFRASTRUCTURE ONLY).

Runs the reference's own CPU implementation (oracle/_ref/libref3dsift.so: the reference sources
compiled unmodified, kind "reference") -- or, when that is absent, the C port (oracle/liboracle.so,
kind "port") -- on the host cores.  The reference is single-threaded and not re-entrant, so
parallelism is P independent worker PROCESSES, one volume each at a time.
Only bench.py's cpu_baseline / --impl reference legs call this.
"""
import ctypes as C
import multiprocessing as mp
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref", "libref3dsift.so")
PORT = os.path.join(HERE, "liboracle.so")


def kind():
    return "reference" if os.path.exists(REF) else "port"


def usable_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:   # leave ~1.5 GB of RAM per worker at MNI size
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    n = max(1, min(n, int(line.split()[1]) // (1536 * 1024)))
    except OSError:
        pass
    return n


_lib = None
_vol = None


def _init(shape_xyz, seed, nblobs, phantom_path):
    global _lib, _vol
    import importlib.util
    spec = importlib.util.spec_from_file_location("s3d_phantom", phantom_path)
    ph = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ph)
    _vol = ph.brain_phantom(tuple(shape_xyz), seed, nblobs)
    os.environ.setdefault("S3D_REF_SCRATCH", "/tmp")
    if os.path.exists(REF):
        _lib = C.CDLL(REF)
        _lib.ref_extract.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    else:
        _lib = C.CDLL(PORT)
        _lib.s3o_extract.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]


def _one(_):
    Z, Y, X = _vol.shape
    p = _vol.ctypes.data_as(C.c_void_p)
    t = time.perf_counter()
    if os.path.exists(REF):
        n = _lib.ref_extract(p, X, Y, Z, 0, 0, None, None, None, None)
    else:
        n = _lib.s3o_extract(p, X, Y, Z, 0, 0, None, None, None, None, None)
    return n, time.perf_counter() - t


class CpuPool:
    """P worker processes, each holding the phantom volume and the CPU library."""

    def __init__(self, shape_xyz, seed, nblobs, phantom_path, procs=None):
        self.procs = procs or usable_cores()
        self.pool = mp.get_context("spawn").Pool(self.procs, initializer=_init,
                                                 initargs=(tuple(shape_xyz), seed, nblobs, phantom_path))
        self.pool.map(_one, range(0))

    def step(self, volumes_per_proc=1):
        """Every worker extracts `volumes_per_proc` volumes; returns (volumes, seconds, rows)."""
        n = self.procs * volumes_per_proc
        t = time.perf_counter()
        res = self.pool.map(_one, range(n), chunksize=volumes_per_proc)
        dt = time.perf_counter() - t
        return n, dt, res[0][0]

    def close(self):
        self.pool.close()
        self.pool.join()
