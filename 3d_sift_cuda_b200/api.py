"""ctypes binding of the C-ABI in include/s3d.h (lib3dsift_b200.so).

Mirrors the reference's interface for the featExtract path: one ``Engine`` per GPU (the reference's
``-dN``), ``Engine.extract(volume, Params(...))`` = ``featExtract [-2+|-2-] [-b|-br|-bn]`` up to the
feature file, and stage-level methods with the meaning of the reference's four CUDA launchers
(SIFT_cuda_Tools.cuh:32-38, 69-76, 202-205, 213-217).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable the calls
raise ``S3DError``.
"""
import ctypes as C
import weakref
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FEATURE_DTYPE = np.dtype([("flag", "<u4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("scale", "<f4"),
                          ("ori", "<f4", (9,)), ("eigs", "<f4", (3,)), ("pc", "<f4", (64,))])
CAND_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("z", "<i4"), ("value", "<f4")])
KEYPOINT_DTYPE = np.dtype([("octave", "<i4"), ("level", "<i4"), ("is_max", "<i4"),
                           ("ix", "<i4"), ("iy", "<i4"), ("iz", "<i4"),
                           ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("scale", "<f4")])

DESC_SIFT, DESC_BRIEF, DESC_RRIEF, DESC_NRRIEF = 0, 1, 2, 3
_STATUS = {0: "ok", 1: "invalid argument", 2: "CUDA error", 3: "out of memory", 4: "capacity exceeded", 5: "unsupported"}

EXPORTS = [
    "s3d_ctx_create", "s3d_ctx_create_on_stream", "s3d_ctx_destroy", "s3d_last_error", "s3d_device_count",
    "s3d_sync", "s3d_stream", "s3d_gaussian_taps", "s3d_blur3d", "s3d_dog", "s3d_subsample2", "s3d_detect",
    "s3d_double_size", "s3d_halve_size", "s3d_extract", "s3d_extract_device", "s3d_extract_host_async",
    "s3d_fetch_features", "s3d_fetch_counts", "s3d_result_device", "s3d_free", "s3d_num_octaves",
    "s3d_get_level", "s3d_get_keypoints", "s3d_get_patches", "s3d_last_launch_count", "s3d_write_features_text",
    "s3d_copy_level_device", "s3d_get_row_keypoints",
    "s3d_batch_create", "s3d_batch_destroy", "s3d_batch_last_error", "s3d_batch_extract", "s3d_batch_extract_device",
    "s3d_batch_launches_per_volume", "s3d_host_alloc", "s3d_host_free",
    "s3d_extract_typed", "s3d_extract_typed_async", "s3d_batch_extract_typed",
    "s3d_write_features_bin", "s3d_read_features_text", "s3d_match", "s3d_match_device",
    "s3d_multi_create", "s3d_multi_destroy", "s3d_multi_last_error", "s3d_multi_gpu_count", "s3d_multi_batch_extract",
    "s3d_multi_extract_slab", "s3d_resample_iso", "s3d_resample_iso_host", "s3d_level_device_ptr",
]

# NIfTI datatype codes accepted by the typed entry points (reference featExtract.cpp:18-77)
DTYPE_CODES = {"uint8": 2, "int16": 4, "int32": 8, "float32": 16, "float64": 64, "int8": 256, "uint16": 512, "uint32": 768}


class S3DError(RuntimeError):
    pass


class _Params(C.Structure):
    _fields_ = [("double_mode", C.c_int), ("descriptor", C.c_int), ("eig_thres", C.c_float),
                ("max_keypoints", C.c_int), ("max_features", C.c_int), ("keep_patches", C.c_int),
                ("input_is_g0", C.c_int), ("octave_base", C.c_int), ("max_octaves", C.c_int), ("slab", C.c_int),
                ("z_off", C.c_int), ("z_global", C.c_int), ("own_z0", C.c_int), ("own_z1", C.c_int),
                ("pre_step_done", C.c_int)]


class Params:
    """featExtract's options (featExtract.cpp:299-350 plus the README's -b/-br/-bn)."""

    def __init__(self, double_mode=0, descriptor=DESC_SIFT, eig_thres=140.0, max_keypoints=0, max_features=0,
                 keep_patches=False, input_is_g0=False, octave_base=0, max_octaves=0, slab=None, pre_step_done=0):
        """``slab`` = (z_off, z_global, own_z0, own_z1) in planes of the octave being run (see include/s3d.h)."""
        z_off, z_global, own_z0, own_z1 = slab if slab is not None else (0, 0, 0, 0)
        self.c = _Params(double_mode, descriptor, eig_thres, max_keypoints, max_features, 1 if keep_patches else 0,
                         1 if input_is_g0 else 0, octave_base, max_octaves, 1 if slab is not None else 0,
                         z_off, z_global, own_z0, own_z1, pre_step_done)


def library_path():
    return os.path.join(_HERE, "lib3dsift_b200.so")


def build_library(verbose=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "all"], check=True, stdout=out)
    return library_path()


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise S3DError("%s is missing: build it with __graft_entry__.build() "
                       "(make -C 3d_sift_cuda_b200/csrc); there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    L.s3d_ctx_create.argtypes = [i, C.POINTER(vp)]
    L.s3d_ctx_create_on_stream.argtypes = [i, vp, C.POINTER(vp)]
    L.s3d_ctx_destroy.argtypes = [vp]
    L.s3d_ctx_destroy.restype = None
    L.s3d_last_error.argtypes = [vp]
    L.s3d_last_error.restype = C.c_char_p
    L.s3d_sync.argtypes = [vp]
    L.s3d_stream.argtypes = [vp]
    L.s3d_stream.restype = vp
    L.s3d_gaussian_taps.argtypes = [f, vp, i]
    L.s3d_blur3d.argtypes = [vp, vp, vp, vp, i, i, i, i, vp, i, vp]
    L.s3d_dog.argtypes = [vp, vp, vp, vp, i, i, i, i]
    for fn in (L.s3d_subsample2, L.s3d_double_size, L.s3d_halve_size):
        fn.argtypes = [vp, vp, i, i, i, i, vp, i]
    L.s3d_detect.argtypes = [vp, vp, vp, i, i, i, i, vp, vp, vp, vp, i]
    L.s3d_extract.argtypes = [vp, vp, i, i, i, vp, C.POINTER(vp), C.POINTER(i)]
    L.s3d_extract_device.argtypes = [vp, vp, i, i, i, vp]
    L.s3d_extract_host_async.argtypes = [vp, vp, i, i, i, vp]
    L.s3d_fetch_features.argtypes = [vp, C.POINTER(vp), C.POINTER(i)]
    L.s3d_fetch_counts.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
    L.s3d_result_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.s3d_free.argtypes = [vp]
    L.s3d_free.restype = None
    L.s3d_num_octaves.argtypes = [vp]
    L.s3d_get_level.argtypes = [vp, i, i, i, vp, C.POINTER(i * 3)]
    L.s3d_get_keypoints.argtypes = [vp, C.POINTER(vp), C.POINTER(i)]
    L.s3d_get_patches.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i)]
    L.s3d_last_launch_count.argtypes = [vp]
    L.s3d_copy_level_device.argtypes = [vp, i, i, i, i, i, vp]
    L.s3d_get_row_keypoints.argtypes = [vp, C.POINTER(vp), C.POINTER(i)]
    L.s3d_write_features_text.argtypes = [C.c_char_p, vp, i, f, i, C.POINTER(C.c_char_p)]
    L.s3d_write_features_bin.argtypes = [C.c_char_p, vp, i, f]
    L.s3d_read_features_text.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(i)]
    L.s3d_batch_create.argtypes = [i, i, C.POINTER(vp)]
    L.s3d_batch_destroy.argtypes = [vp]
    L.s3d_batch_destroy.restype = None
    L.s3d_batch_last_error.argtypes = [vp]
    L.s3d_batch_last_error.restype = C.c_char_p
    L.s3d_batch_extract.argtypes = [vp, C.POINTER(vp), i, i, i, i, vp, C.POINTER(vp), C.POINTER(i)]
    L.s3d_batch_extract_device.argtypes = [vp, C.POINTER(vp), i, i, i, i, vp, C.POINTER(i), C.POINTER(i)]
    L.s3d_batch_launches_per_volume.argtypes = [vp]
    L.s3d_extract_typed.argtypes = [vp, vp, i, i, i, i, vp, C.POINTER(vp), C.POINTER(i)]
    L.s3d_extract_typed_async.argtypes = [vp, vp, i, i, i, i, vp]
    L.s3d_batch_extract_typed.argtypes = [vp, C.POINTER(vp), i, i, i, i, i, vp, C.POINTER(vp), C.POINTER(i)]
    L.s3d_match.argtypes = [vp, vp, i, vp, i, i, vp, vp]
    L.s3d_match_device.argtypes = [vp, vp, i, vp, i, i, vp, vp]
    L.s3d_multi_create.argtypes = [i, vp, i, C.POINTER(vp)]
    L.s3d_multi_destroy.argtypes = [vp]
    L.s3d_multi_destroy.restype = None
    L.s3d_multi_last_error.argtypes = [vp]
    L.s3d_multi_last_error.restype = C.c_char_p
    L.s3d_multi_gpu_count.argtypes = [vp]
    L.s3d_multi_batch_extract.argtypes = [vp, C.POINTER(vp), i, i, i, i, vp, C.POINTER(vp), C.POINTER(i)]
    L.s3d_multi_extract_slab.argtypes = [vp, vp, i, i, i, vp, C.POINTER(vp), C.POINTER(i)]
    L.s3d_resample_iso.argtypes = [vp, vp, i, i, i, i, vp, i, i, i, i, f, f, f]
    L.s3d_resample_iso_host.argtypes = [vp, vp, i, i, i, vp, i, i, i, f, f, f]
    L.s3d_level_device_ptr.argtypes = [vp, i, i, i, C.POINTER(vp), C.POINTER(i), C.POINTER(i * 3)]
    L.s3d_host_alloc.argtypes = [C.c_size_t]
    L.s3d_host_alloc.restype = vp
    L.s3d_host_free.argtypes = [vp]
    L.s3d_host_free.restype = None
    _LIB = L
    return L


def gaussian_taps(sigma):
    """Normalised taps for one blur (host arithmetic, GaussianMask.cpp:12-57, 241-265)."""
    L = load_library()
    buf = np.zeros(129, np.float32)
    n = L.s3d_gaussian_taps(sigma, buf.ctypes.data_as(C.c_void_p), 129)
    if n <= 0:
        raise S3DError("sigma %g needs %d taps" % (sigma, -n))
    return buf[:n].copy()


def _copy_out(ptr, count, dtype, free):
    if not ptr or count == 0:
        if ptr:
            free(ptr)
        return np.zeros(0, dtype)
    # zero copy: the array views the malloc'ed C buffer, which is released (s3d_free) when the array is collected
    nbytes = count * np.dtype(dtype).itemsize
    addr = ptr.value if hasattr(ptr, "value") else int(ptr)
    buf = (C.c_char * nbytes).from_address(addr)
    weakref.finalize(buf, free, C.c_void_p(addr))
    return np.frombuffer(buf, dtype=dtype)


def _dptr(t):
    """Device pointer of a torch CUDA tensor (float32 / int32, contiguous)."""
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


class Engine:
    """One GPU's extraction context (reference: FEATUREIO.device / -dN)."""

    def __init__(self, device=0, stream=None):
        self.L = load_library()
        self.ctx = C.c_void_p()
        if stream is None:
            st = self.L.s3d_ctx_create(device, C.byref(self.ctx))
        else:
            st = self.L.s3d_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(self.ctx))
        if st != 0:
            msg = self.L.s3d_last_error(self.ctx).decode() if self.ctx else "no usable CUDA device"
            if self.ctx:
                self.L.s3d_ctx_destroy(self.ctx)
                self.ctx = C.c_void_p()
            raise S3DError("s3d_ctx_create(%d): %s (%s)" % (device, _STATUS.get(st, st), msg))
        self.device = device

    def close(self):
        if getattr(self, "ctx", None):
            self.L.s3d_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st, what):
        if st != 0:
            raise S3DError("%s: %s (%s)" % (what, _STATUS.get(st, st), self.L.s3d_last_error(self.ctx).decode()))

    @property
    def stream(self):
        return self.L.s3d_stream(self.ctx)

    def sync(self):
        self._ck(self.L.s3d_sync(self.ctx), "s3d_sync")

    # ---- pipeline level ----------------------------------------------------------------------------
    def extract(self, volume, params=None):
        """Host volume (numpy (Z, Y, X) float32) -> structured array of feature rows."""
        params = params or Params()
        v = np.ascontiguousarray(volume, dtype=np.float32)
        Z, Y, X = v.shape
        out, n = C.c_void_p(), C.c_int()
        self._ck(self.L.s3d_extract(self.ctx, v.ctypes.data_as(C.c_void_p), X, Y, Z, C.byref(params.c),
                                    C.byref(out), C.byref(n)), "s3d_extract")
        return _copy_out(out, n.value, FEATURE_DTYPE, self.L.s3d_free)

    def extract_typed(self, volume, params=None):
        """Host volume of any NIfTI scalar dtype (numpy (Z, Y, X)): raw voxels go over PCIe, the cast to float
        runs on the device (s3d_extract_typed) -- same rows as extract(volume.astype(float32))."""
        params = params or Params()
        v = np.ascontiguousarray(volume)
        code = DTYPE_CODES[v.dtype.name]
        Z, Y, X = v.shape
        out, n = C.c_void_p(), C.c_int()
        self._ck(self.L.s3d_extract_typed(self.ctx, v.ctypes.data_as(C.c_void_p), code, X, Y, Z, C.byref(params.c),
                                          C.byref(out), C.byref(n)), "s3d_extract_typed")
        return _copy_out(out, n.value, FEATURE_DTYPE, self.L.s3d_free)

    def extract_device(self, d_volume, shape_xyz, params=None):
        """Dense device volume (torch tensor or raw pointer); enqueues only."""
        params = params or Params()
        self._last_params = params
        X, Y, Z = shape_xyz
        p = _dptr(d_volume) if hasattr(d_volume, "data_ptr") else C.c_void_p(d_volume)
        self._ck(self.L.s3d_extract_device(self.ctx, p, X, Y, Z, C.byref(params.c)), "s3d_extract_device")

    def extract_host_async(self, h_volume, params=None):
        """Host volume (numpy array or pinned torch tensor, (Z, Y, X) float32); enqueues H2D + the path."""
        params = params or Params()
        self._last_params = params
        if hasattr(h_volume, "data_ptr"):
            Z, Y, X = h_volume.shape
            p = C.c_void_p(h_volume.data_ptr())
        else:
            assert h_volume.dtype == np.float32 and h_volume.flags["C_CONTIGUOUS"]
            Z, Y, X = h_volume.shape
            p = h_volume.ctypes.data_as(C.c_void_p)
        self._ck(self.L.s3d_extract_host_async(self.ctx, p, X, Y, Z, C.byref(params.c)), "s3d_extract_host_async")

    def fetch_features(self):
        out, n = C.c_void_p(), C.c_int()
        self._ck(self.L.s3d_fetch_features(self.ctx, C.byref(out), C.byref(n)), "s3d_fetch_features")
        return _copy_out(out, n.value, FEATURE_DTYPE, self.L.s3d_free)

    def fetch_counts(self):
        nk, nf = C.c_int(), C.c_int()
        self._ck(self.L.s3d_fetch_counts(self.ctx, C.byref(nk), C.byref(nf)), "s3d_fetch_counts")
        return nk.value, nf.value

    def num_octaves(self):
        return self.L.s3d_num_octaves(self.ctx)

    def level(self, octave, level, dog=False):
        dims = (C.c_int * 3)()
        self._ck(self.L.s3d_get_level(self.ctx, octave, 1 if dog else 0, level, None, C.byref(dims)), "s3d_get_level")
        X, Y, Z = dims[0], dims[1], dims[2]
        out = np.empty((Z, Y, X), np.float32)
        self._ck(self.L.s3d_get_level(self.ctx, octave, 1 if dog else 0, level, out.ctypes.data_as(C.c_void_p),
                                      C.byref(dims)), "s3d_get_level")
        return out

    def keypoints(self):
        out, n = C.c_void_p(), C.c_int()
        self._ck(self.L.s3d_get_keypoints(self.ctx, C.byref(out), C.byref(n)), "s3d_get_keypoints")
        return _copy_out(out, n.value, KEYPOINT_DTYPE, self.L.s3d_free)

    def patches(self):
        pa, pr, n = C.c_void_p(), C.c_void_p(), C.c_int()
        self._ck(self.L.s3d_get_patches(self.ctx, C.byref(pa), C.byref(pr), C.byref(n)), "s3d_get_patches")
        nn = n.value
        a = _copy_out(pa, nn * 1331, np.float32, self.L.s3d_free).reshape(nn, 11, 11, 11)
        b = _copy_out(pr, nn * 64, np.float32, self.L.s3d_free).reshape(nn, 64)
        return a, b

    def launch_count(self):
        return self.L.s3d_last_launch_count(self.ctx)

    def copy_level_device(self, octave, level, z0, z1, d_dst, dog=False):
        """Planes [z0, z1) of a level of the last extraction -> dense device tensor (z1-z0, Y, X)."""
        self._ck(self.L.s3d_copy_level_device(self.ctx, octave, 1 if dog else 0, level, z0, z1, _dptr(d_dst)),
                 "s3d_copy_level_device")

    def row_keypoints(self):
        """For every feature row of the last extraction, the index of its keypoint."""
        out, n = C.c_void_p(), C.c_int()
        self._ck(self.L.s3d_get_row_keypoints(self.ctx, C.byref(out), C.byref(n)), "s3d_get_row_keypoints")
        return _copy_out(out, n.value, np.int32, self.L.s3d_free)

    # ---- stage level (torch CUDA tensors shaped (Z, Y, pitch), float32) ------------------------------
    def blur3d(self, d_in, d_tmp, d_out, X, taps, d_dog=None):
        Z, Y, pitch = d_in.shape
        t = np.ascontiguousarray(taps, np.float32)
        self._ck(self.L.s3d_blur3d(self.ctx, _dptr(d_in), _dptr(d_tmp), _dptr(d_out), X, Y, Z, pitch,
                                   t.ctypes.data_as(C.c_void_p), len(t), _dptr(d_dog) if d_dog is not None else None),
                 "s3d_blur3d")

    def dog(self, d_a, d_b, d_out, X):
        Z, Y, pitch = d_a.shape
        self._ck(self.L.s3d_dog(self.ctx, _dptr(d_a), _dptr(d_b), _dptr(d_out), X, Y, Z, pitch), "s3d_dog")

    def _resize(self, fn, name, d_in, X, d_out):
        Z, Y, pitch = d_in.shape
        self._ck(fn(self.ctx, _dptr(d_in), X, Y, Z, pitch, _dptr(d_out), d_out.shape[2]), name)

    def subsample2(self, d_in, X, d_out):
        self._resize(self.L.s3d_subsample2, "s3d_subsample2", d_in, X, d_out)

    def double_size(self, d_in, X, d_out):
        self._resize(self.L.s3d_double_size, "s3d_double_size", d_in, X, d_out)

    def halve_size(self, d_in, X, d_out):
        self._resize(self.L.s3d_halve_size, "s3d_halve_size", d_in, X, d_out)

    def match(self, feats_a, feats_b, k=2):
        """Exact k nearest neighbours of every row of feats_a among feats_b on the reference's descriptor distance
        (s3d_match): returns (idx [nA, k] int32, dist [nA, k] float32), neighbours in (distance, index) order."""
        a = np.ascontiguousarray(feats_a, dtype=FEATURE_DTYPE)
        b = np.ascontiguousarray(feats_b, dtype=FEATURE_DTYPE)
        idx = np.empty((len(a), k), np.int32)
        dist = np.empty((len(a), k), np.float32)
        self._ck(self.L.s3d_match(self.ctx, a.ctypes.data_as(C.c_void_p), len(a), b.ctypes.data_as(C.c_void_p), len(b), k,
                                  idx.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p)), "s3d_match")
        return idx, dist

    def detect(self, d_finer, d_centre, X, cap=1 << 16):
        """Returns (minima, maxima) as CAND_DTYPE arrays in raster order."""
        import torch
        Z, Y, pitch = d_centre.shape
        dev = d_centre.device
        mins = torch.zeros(cap * 4, dtype=torch.int32, device=dev)
        maxs = torch.zeros(cap * 4, dtype=torch.int32, device=dev)
        cnt = torch.zeros(2, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)   # the engine has its own stream
        self._ck(self.L.s3d_detect(self.ctx, _dptr(d_finer), _dptr(d_centre), X, Y, Z, pitch,
                                   _dptr(mins), C.c_void_p(cnt.data_ptr()), _dptr(maxs), C.c_void_p(cnt.data_ptr() + 4), cap),
                 "s3d_detect")
        self.sync()
        n = cnt.cpu().numpy()
        if n[0] > cap or n[1] > cap:
            raise S3DError("s3d_detect: %d/%d candidates exceed cap %d" % (n[0], n[1], cap))
        a = mins.cpu().numpy().view(CAND_DTYPE)[:n[0]].copy()
        b = maxs.cpu().numpy().view(CAND_DTYPE)[:n[1]].copy()
        return a, b


class Batch:
    """Several extraction contexts on one GPU fed round-robin (s3d_batch, BASELINE.json config 4 per GPU)."""

    def __init__(self, device=0, n_contexts=4):
        self.L = load_library()
        self.b = C.c_void_p()
        st = self.L.s3d_batch_create(device, n_contexts, C.byref(self.b))
        if st != 0:
            raise S3DError("s3d_batch_create(%d, %d): %s" % (device, n_contexts, _STATUS.get(st, st)))
        self.n_contexts = n_contexts

    def close(self):
        if getattr(self, "b", None):
            self.L.s3d_batch_destroy(self.b)
            self.b = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st, what):
        if st != 0:
            raise S3DError("%s: %s (%s)" % (what, _STATUS.get(st, st), self.L.s3d_batch_last_error(self.b).decode()))

    @staticmethod
    def _ptrs(vols):
        arr = (C.c_void_p * len(vols))()
        for k, v in enumerate(vols):
            arr[k] = v.data_ptr() if hasattr(v, "data_ptr") else v.ctypes.data
        return arr

    @staticmethod
    def _checked(vols, dtype=None):
        """Same shape everywhere, C-contiguous; numpy inputs are converted if needed (the returned list keeps
        the converted arrays alive for the duration of the call)."""
        out = []
        shape = tuple(vols[0].shape)
        for v in vols:
            if tuple(v.shape) != shape or len(shape) != 3:
                raise S3DError("batch volumes must all have the same (Z, Y, X) shape")
            if hasattr(v, "data_ptr"):
                if not v.is_contiguous() or (dtype is not None and str(v.dtype).replace("torch.", "") != dtype):
                    raise S3DError("torch volumes must be contiguous%s" % (" " + dtype if dtype else ""))
                out.append(v)
            else:
                out.append(np.ascontiguousarray(v, dtype=dtype) if dtype else np.ascontiguousarray(v))
        return out

    def extract(self, h_volumes, params=None):
        """List of host volumes ((Z, Y, X) float32 numpy arrays or pinned torch tensors of one shape) ->
        list of feature-row arrays, in input order."""
        params = params or Params()
        n = len(h_volumes)
        if n == 0:
            return []
        h_volumes = self._checked(h_volumes, "float32")
        Z, Y, X = h_volumes[0].shape
        rows = (C.c_void_p * n)()
        cnt = (C.c_int * n)()
        st = self.L.s3d_batch_extract(self.b, self._ptrs(h_volumes), n, X, Y, Z, C.byref(params.c), rows, cnt)
        # wrap (and thereby own) whatever was malloc'ed before a failing volume stopped the batch, then report
        out = [_copy_out(C.c_void_p(rows[k]), cnt[k], FEATURE_DTYPE, self.L.s3d_free) if rows[k] else np.zeros(0, FEATURE_DTYPE) for k in range(n)]
        self._ck(st, "s3d_batch_extract")
        return out

    def extract_typed(self, h_volumes, params=None):
        """Like extract() for host volumes of one NIfTI scalar dtype (numpy arrays or pinned torch tensors)."""
        params = params or Params()
        n = len(h_volumes)
        if n == 0:
            return []
        v0 = h_volumes[0]
        name = str(v0.dtype).replace("torch.", "")
        if name not in DTYPE_CODES:
            raise S3DError("unsupported voxel datatype %s" % name)
        h_volumes = self._checked(h_volumes, name)
        code = DTYPE_CODES[name]
        Z, Y, X = v0.shape
        rows = (C.c_void_p * n)()
        cnt = (C.c_int * n)()
        st = self.L.s3d_batch_extract_typed(self.b, self._ptrs(h_volumes), code, n, X, Y, Z, C.byref(params.c), rows, cnt)
        out = [_copy_out(C.c_void_p(rows[k]), cnt[k], FEATURE_DTYPE, self.L.s3d_free) if rows[k] else np.zeros(0, FEATURE_DTYPE) for k in range(n)]
        self._ck(st, "s3d_batch_extract_typed")
        return out

    def extract_device(self, d_volumes, shape_xyz, params=None):
        """List of dense device volumes -> (n_keypoints, n_rows) per volume; rows stay on the device."""
        params = params or Params()
        n = len(d_volumes)
        X, Y, Z = shape_xyz
        nk = (C.c_int * n)()
        nr = (C.c_int * n)()
        self._ck(self.L.s3d_batch_extract_device(self.b, self._ptrs(d_volumes), n, X, Y, Z, C.byref(params.c), nk, nr),
                 "s3d_batch_extract_device")
        return list(nk), list(nr)

    def launches_per_volume(self):
        return self.L.s3d_batch_launches_per_volume(self.b)


class Multi:
    """Several GPUs of one box behind one handle (s3d_multi_*): batch sharding and z-slab decomposition from a
    single process, one host thread per GPU inside the library.  ``devices`` may list a device more than once."""

    def __init__(self, devices, contexts_per_gpu=4):
        self.L = load_library()
        self.h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        st = self.L.s3d_multi_create(len(devices), devs, contexts_per_gpu, C.byref(self.h))
        if st != 0:
            msg = self.L.s3d_multi_last_error(self.h).decode() if self.h else ""
            raise S3DError("s3d_multi_create failed (status %d): %s" % (st, msg))

    def close(self):
        if self.h:
            self.L.s3d_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st, what):
        if st != 0:
            raise S3DError("%s failed (status %d): %s" % (what, st, self.L.s3d_multi_last_error(self.h).decode()))

    def extract_slab(self, volume, params=None):
        """One volume split into z slabs over the handle's GPUs; rows in the reference's order."""
        p = (params or Params()).c
        v = np.ascontiguousarray(volume, np.float32)
        Z, Y, X = v.shape
        out, n = C.c_void_p(), C.c_int()
        self._ck(self.L.s3d_multi_extract_slab(self.h, v.ctypes.data_as(C.c_void_p), X, Y, Z, C.byref(p), C.byref(out), C.byref(n)),
                 "s3d_multi_extract_slab")
        return _copy_out(out, n.value, FEATURE_DTYPE, self.L.s3d_free)

    def batch_extract(self, volumes, params=None):
        """Volume i on GPU i mod n; list of row arrays in input order."""
        p = (params or Params()).c
        vols = [np.ascontiguousarray(v, np.float32) for v in volumes]
        if not vols:
            return []
        Z, Y, X = vols[0].shape
        assert all(v.shape == vols[0].shape for v in vols), "volumes of one batch share a shape"
        ptrs = (C.c_void_p * len(vols))(*[v.ctypes.data_as(C.c_void_p) for v in vols])
        rows = (C.c_void_p * len(vols))()
        n_rows = (C.c_int * len(vols))()
        st = self.L.s3d_multi_batch_extract(self.h, ptrs, len(vols), X, Y, Z, C.byref(p), rows, n_rows)
        out = [_copy_out(C.c_void_p(rows[k]), n_rows[k], FEATURE_DTYPE, self.L.s3d_free) if rows[k] else None for k in range(len(vols))]
        self._ck(st, "s3d_multi_batch_extract")
        return out


def write_features_bin(path, feats, eig_thres=-1.0):
    """msFeature3DVectorOutputBin (MultiScale.h:228-303)."""
    L = load_library()
    f = np.ascontiguousarray(feats, dtype=FEATURE_DTYPE)
    st = L.s3d_write_features_bin(path.encode(), f.ctypes.data_as(C.c_void_p), len(f), eig_thres)
    if st != 0:
        raise S3DError("s3d_write_features_bin: %s" % _STATUS.get(st, st))


def read_features_text(path):
    """msFeature3DVectorInputText (MultiScale.h:305-384): feature rows of a text .key file."""
    L = load_library()
    out, n = C.c_void_p(), C.c_int()
    st = L.s3d_read_features_text(path.encode(), C.byref(out), C.byref(n))
    if st != 0:
        raise S3DError("s3d_read_features_text(%s): %s" % (path, _STATUS.get(st, st)))
    return _copy_out(out, n.value, FEATURE_DTYPE, L.s3d_free)


def write_features_text(path, feats, shape_xyz, eig_thres=140.0, comments=None):
    """The reference's text feature file (MultiScale.h:386-474, featExtract.cpp:542-575)."""
    L = load_library()
    f = np.ascontiguousarray(feats, dtype=FEATURE_DTYPE)
    if comments is None:
        comments = [
            "Extraction Voxel Resolution (ijk) : %d %d %d" % tuple(shape_xyz),
            "Extraction Voxel Size (mm)  (ijk) : %f %f %f" % (1.0, 1.0, 1.0),
            "Feature Coordinate Space: voxels: 1.0 0.0 0.0 0.0 0.0 1.0 0.0 0.0 0.0 0.0 1.0 0.0 0.0 0.0 0.0 1.0",
        ]
    arr = (C.c_char_p * len(comments))(*[c.encode() for c in comments])
    st = L.s3d_write_features_text(path.encode(), f.ctypes.data_as(C.c_void_p), len(f), eig_thres, len(comments), arr)
    if st != 0:
        raise S3DError("s3d_write_features_text(%s): %s" % (path, _STATUS.get(st, st)))
