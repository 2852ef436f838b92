"""ctypes bindings for the two CHECKERS used by the tests (never by the product path):

* ``Oracle``    -- oracle/liboracle.so, the plain-C restatement (oracle/sift3d_oracle.c);
* ``Reference`` -- oracle/_ref/libref3dsift.so, the reference's own sources behind
                   oracle/ref_driver.cpp (present when built in a container that has
                   /root/reference; it travels to the GPU box prebuilt).

Both expose the same method names and return numpy arrays, so tests can run the same
assertions against either.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class Feature(C.Structure):
    """Feature3DInfo layout (reference MultiScale.h:111-129)."""
    _fields_ = [("flag", C.c_uint), ("x", C.c_float), ("y", C.c_float), ("z", C.c_float),
                ("scale", C.c_float), ("ori", C.c_float * 9), ("eigs", C.c_float * 3),
                ("pc", C.c_float * 64)]


FEATURE_DTYPE = np.dtype([("flag", "<u4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("scale", "<f4"),
                          ("ori", "<f4", (9,)), ("eigs", "<f4", (3,)), ("pc", "<f4", (64,))])
CAND_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("z", "<i4"), ("value", "<f4")])
KEYPOINT_DTYPE = np.dtype([("octave", "<i4"), ("level", "<i4"), ("is_max", "<i4"),
                           ("ix", "<i4"), ("iy", "<i4"), ("iz", "<i4"),
                           ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("scale", "<f4")])
assert FEATURE_DTYPE.itemsize == C.sizeof(Feature) == 324


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


def build_oracle():
    """(Re)build oracle/liboracle.so and, when the reference is present, oracle/_ref."""
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.DEVNULL)


def _take(ptr, n, dtype, free):
    """Copy n records out of a malloc'ed block and free it."""
    if not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = n * np.dtype(dtype).itemsize
    buf = (C.c_char * nbytes).from_address(ptr.value if isinstance(ptr, C.c_void_p) else ptr)
    out = np.frombuffer(bytes(buf), dtype=dtype).copy()
    free(ptr)
    return out


class Oracle:
    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = L = C.CDLL(path)
        L.s3o_gaussian_taps.argtypes = [C.c_float, C.c_void_p, C.c_int]
        L.s3o_blur3d.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float]
        L.s3o_blur3d_taps.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.s3o_dog.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        for f in (L.s3o_subsample, L.s3o_double_size, L.s3o_halve_size):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.s3o_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.s3o_octave_levels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s3o_extract.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s3o_free.argtypes = [C.c_void_p]
        L.s3o_sample_patch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_void_p, C.c_void_p]
        for f in (L.s3o_normalize_patch, L.s3o_rank):
            f.argtypes = [C.c_void_p]
        L.s3o_eigen_orientation.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.s3o_canonical_orientations.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.s3o_descriptor_sift.argtypes = [C.c_void_p, C.c_void_p]
        L.s3o_descriptor_brief.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.s3o_knn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]

    def knn(self, pcs_a, pcs_b, k):
        """Exhaustive kNN on DistSqrPCs; pcs_*: [n, 64] float32 descriptors."""
        a = np.ascontiguousarray(pcs_a, np.float32); b = np.ascontiguousarray(pcs_b, np.float32)
        idx = np.empty((len(a), k), np.int32); dist = np.empty((len(a), k), np.float32)
        self.lib.s3o_knn(a.ctypes.data_as(C.c_void_p), len(a), b.ctypes.data_as(C.c_void_p), len(b), k,
                         idx.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p))
        return idx, dist

    # --- voxel stages (vol: numpy (Z, Y, X) float32) ---
    def taps(self, sigma):
        buf = np.zeros(129, np.float32)
        n = self.lib.s3o_gaussian_taps(sigma, buf.ctypes.data_as(C.c_void_p), 129)
        assert n > 0
        return buf[:n].copy()

    def blur(self, vol, sigma):
        v, p = _f32(vol)
        out = np.empty_like(v)
        Z, Y, X = v.shape
        assert self.lib.s3o_blur3d(p, out.ctypes.data_as(C.c_void_p), X, Y, Z, sigma) == 1
        return out

    def dog(self, a, b):
        a, pa = _f32(a)
        b, pb = _f32(b)
        out = np.empty_like(a)
        self.lib.s3o_dog(pa, pb, out.ctypes.data_as(C.c_void_p), a.size)
        return out

    def _resize(self, fn, vol, num, den):
        v, p = _f32(vol)
        Z, Y, X = v.shape
        out = np.empty((Z * num // den, Y * num // den, X * num // den), np.float32)
        fn(p, out.ctypes.data_as(C.c_void_p), X, Y, Z)
        return out

    def subsample(self, vol):
        return self._resize(self.lib.s3o_subsample, vol, 1, 2)

    def double_size(self, vol):
        return self._resize(self.lib.s3o_double_size, vol, 2, 1)

    def halve_size(self, vol):
        return self._resize(self.lib.s3o_halve_size, vol, 1, 2)

    def detect(self, finer, centre, cap=1 << 20):
        f, pf = _f32(finer)
        c, pc = _f32(centre)
        Z, Y, X = c.shape
        mins = np.zeros(cap, CAND_DTYPE)
        maxs = np.zeros(cap, CAND_DTYPE)
        nmin, nmax = C.c_int(), C.c_int()
        self.lib.s3o_detect(pf, pc, X, Y, Z, mins.ctypes.data_as(C.c_void_p), C.byref(nmin),
                            maxs.ctypes.data_as(C.c_void_p), C.byref(nmax), cap)
        return mins[:nmin.value].copy(), maxs[:nmax.value].copy()

    def octave_levels(self, g0):
        g0, p = _f32(g0)
        Z, Y, X = g0.shape
        g = np.empty((6, Z, Y, X), np.float32)
        d = np.empty((5, Z, Y, X), np.float32)
        sig = np.zeros(6, np.float32)
        gp = (C.c_void_p * 6)(*[g[i].ctypes.data for i in range(6)])
        dp = (C.c_void_p * 5)(*[d[i].ctypes.data for i in range(5)])
        self.lib.s3o_octave_levels(p, X, Y, Z, gp, dp, sig.ctypes.data_as(C.c_void_p))
        return g, d, sig

    # --- whole path ---
    def extract(self, vol, double_mode=0, descriptor=0, want_keypoints=False):
        v, p = _f32(vol)
        Z, Y, X = v.shape
        feats, patches, prerank, kps = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nkp = C.c_int()
        n = self.lib.s3o_extract(p, X, Y, Z, double_mode, descriptor, C.byref(feats), C.byref(patches),
                                 C.byref(prerank), C.byref(kps) if want_keypoints else None, C.byref(nkp))
        assert n >= 0
        out = {
            "features": _take(feats, n, FEATURE_DTYPE, self.lib.s3o_free),
            "patches": _take(patches, n * 1331, np.float32, self.lib.s3o_free).reshape(n, 11, 11, 11),
            "prerank": _take(prerank, n * 64, np.float32, self.lib.s3o_free).reshape(n, 64),
        }
        if want_keypoints:
            out["keypoints"] = _take(kps, nkp.value, KEYPOINT_DTYPE, self.lib.s3o_free)
        return out


class Reference:
    """The reference's own code (oracle/_ref/libref3dsift.so)."""

    @staticmethod
    def path():
        return os.path.join(ORACLE_DIR, "_ref", "libref3dsift.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.path())

    def __init__(self):
        self.lib = L = C.CDLL(self.path())
        L.ref_gaussian_taps.argtypes = [C.c_float, C.c_void_p, C.c_int]
        L.ref_blur3d.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float]
        L.ref_dog.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        for f in (L.ref_subsample, L.ref_double_size, L.ref_halve_size):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_extract.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_write_text.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int]
        if hasattr(L, "ref_write_bin"):
            L.ref_write_bin.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_float]
            L.ref_read_text.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ref_free.argtypes = [C.c_void_p]
        if hasattr(L, "ref_knn"):
            L.ref_knn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]

    def taps(self, sigma):
        buf = np.zeros(129, np.float32)
        n = self.lib.ref_gaussian_taps(sigma, buf.ctypes.data_as(C.c_void_p), 129)
        assert n > 0
        return buf[:n].copy()

    def blur(self, vol, sigma):
        v, p = _f32(vol)
        out = np.empty_like(v)
        Z, Y, X = v.shape
        assert self.lib.ref_blur3d(p, out.ctypes.data_as(C.c_void_p), X, Y, Z, sigma) == 1
        return out

    def dog(self, a, b):
        a, pa = _f32(a)
        b, pb = _f32(b)
        out = np.empty_like(a)
        Z, Y, X = a.shape
        self.lib.ref_dog(pa, pb, out.ctypes.data_as(C.c_void_p), X, Y, Z)
        return out

    def _resize(self, fn, vol, num, den):
        v, p = _f32(vol)
        Z, Y, X = v.shape
        out = np.empty((Z * num // den, Y * num // den, X * num // den), np.float32)
        fn(p, out.ctypes.data_as(C.c_void_p), X, Y, Z)
        return out

    def subsample(self, vol):
        return self._resize(self.lib.ref_subsample, vol, 1, 2)

    def double_size(self, vol):
        return self._resize(self.lib.ref_double_size, vol, 2, 1)

    def halve_size(self, vol):
        return self._resize(self.lib.ref_halve_size, vol, 1, 2)

    def detect(self, finer, centre, cap=1 << 20):
        f, pf = _f32(finer)
        c, pc = _f32(centre)
        Z, Y, X = c.shape
        mn_xyz, mx_xyz = np.zeros((cap, 3), np.int32), np.zeros((cap, 3), np.int32)
        mn_v, mx_v = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        nmin, nmax = C.c_int(), C.c_int()
        self.lib.ref_detect(pf, pc, X, Y, Z, mn_xyz.ctypes.data_as(C.c_void_p), mn_v.ctypes.data_as(C.c_void_p),
                            C.byref(nmin), mx_xyz.ctypes.data_as(C.c_void_p), mx_v.ctypes.data_as(C.c_void_p),
                            C.byref(nmax), cap)

        def pack(xyz, v, n):
            out = np.zeros(n, CAND_DTYPE)
            out["x"], out["y"], out["z"], out["value"] = xyz[:n, 0], xyz[:n, 1], xyz[:n, 2], v[:n]
            return out
        return pack(mn_xyz, mn_v, nmin.value), pack(mx_xyz, mx_v, nmax.value)

    def extract(self, vol, double_mode=0, descriptor=0):
        v, p = _f32(vol)
        Z, Y, X = v.shape
        feats, patches, prerank = C.c_void_p(), C.c_void_p(), C.c_void_p()
        sec = C.c_double()
        n = self.lib.ref_extract(p, X, Y, Z, double_mode, descriptor, C.byref(feats), C.byref(patches),
                                 C.byref(prerank), C.byref(sec))
        assert n >= 0
        return {
            "features": _take(feats, n, FEATURE_DTYPE, self.lib.ref_free),
            "patches": _take(patches, n * 1331, np.float32, self.lib.ref_free).reshape(n, 11, 11, 11),
            "prerank": _take(prerank, n * 64, np.float32, self.lib.ref_free).reshape(n, 64),
            "seconds": sec.value,
        }

    def knn(self, feats_a, feats_b, k):
        """Exhaustive kNN with the reference's own Feature3DInfo::DistSqrPCs; feats_*: Feature records."""
        a = np.ascontiguousarray(feats_a); b = np.ascontiguousarray(feats_b)
        idx = np.empty((len(a), k), np.int32); dist = np.empty((len(a), k), np.float32)
        self.lib.ref_knn(a.ctypes.data_as(C.c_void_p), len(a), b.ctypes.data_as(C.c_void_p), len(b), k,
                         idx.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p))
        return idx, dist

    def write_text(self, feats, path, shape_xyz):
        f = np.ascontiguousarray(feats, dtype=FEATURE_DTYPE)
        return self.lib.ref_write_text(f.ctypes.data_as(C.c_void_p), len(f), path.encode(), *shape_xyz)

    def write_bin(self, feats, path, eig_thres=-1.0):
        f = np.ascontiguousarray(feats, dtype=FEATURE_DTYPE)
        return self.lib.ref_write_bin(f.ctypes.data_as(C.c_void_p), len(f), path.encode(), eig_thres)

    def read_text(self, path):
        out = C.c_void_p()
        n = self.lib.ref_read_text(path.encode(), C.byref(out))
        assert n >= 0
        return _take(out, n, FEATURE_DTYPE, self.lib.ref_free)
