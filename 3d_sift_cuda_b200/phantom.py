"""Synthetic phantom volumes: fp32, x fastest, numpy shape (Z, Y, X).

``blob_phantom``  -- BASELINE.json config 1: constant background 50 + K Gaussian blobs.
``brain_phantom`` -- configs 2-5: ellipsoid "head" of intensity 100 + K blobs masked by the head.

Deterministic in (shape, seed, K): ``numpy.random.default_rng(seed)`` with scalar draws in the
order cx, cy, cz, sigma, a per blob.
"""
import numpy as np


def _add_blob(vol, cx, cy, cz, sigma, a, mask=None):
    Z, Y, X = vol.shape
    h = int(4 * sigma) + 1
    x0, x1 = max(0, int(cx) - h), min(X, int(cx) + h + 1)
    y0, y1 = max(0, int(cy) - h), min(Y, int(cy) + h + 1)
    z0, z1 = max(0, int(cz) - h), min(Z, int(cz) + h + 1)
    if x0 >= x1 or y0 >= y1 or z0 >= z1:
        return
    zz, yy, xx = np.meshgrid(np.arange(z0, z1), np.arange(y0, y1), np.arange(x0, x1), indexing="ij")
    r2 = (xx - cx) ** 2 + (yy - cy) ** 2 + (zz - cz) ** 2
    blob = a * np.exp(-r2 / (2.0 * sigma * sigma))
    if mask is not None:
        blob = blob * mask[z0:z1, y0:y1, x0:x1]
    vol[z0:z1, y0:y1, x0:x1] += blob


def blob_phantom(shape_xyz=(128, 128, 128), seed=2, nblobs=200):
    """Config 1 phantom: background 50, centres U[12, dim-12], sigma U[1.5,6], a 60*U[-1,1]."""
    X, Y, Z = shape_xyz
    rng = np.random.default_rng(seed)
    vol = np.full((Z, Y, X), 50.0, dtype=np.float64)
    mx, my, mz = min(12, X // 4), min(12, Y // 4), min(12, Z // 4)   # 12 at the named sizes
    for _ in range(nblobs):
        cx = rng.uniform(mx, X - mx)
        cy = rng.uniform(my, Y - my)
        cz = rng.uniform(mz, Z - mz)
        sigma = rng.uniform(1.5, 6.0)
        a = 60.0 * rng.uniform(-1.0, 1.0)
        _add_blob(vol, cx, cy, cz, sigma, a)
    return np.ascontiguousarray(vol.astype(np.float32))


def brain_phantom(shape_xyz=(182, 218, 182), seed=1, nblobs=400):
    """Configs 2-5 phantom: ellipsoid head (semi-axes 70,90,70 at MNI size, scaled otherwise)."""
    X, Y, Z = shape_xyz
    rng = np.random.default_rng(seed)
    cx0, cy0, cz0 = X / 2.0, Y / 2.0, Z / 2.0
    ax, ay, az = 70.0 * X / 182.0, 90.0 * Y / 218.0, 70.0 * Z / 182.0
    z, y, x = np.ogrid[0:Z, 0:Y, 0:X]
    mask = (((x - cx0) / ax) ** 2 + ((y - cy0) / ay) ** 2 + ((z - cz0) / az) ** 2) <= 1.0
    vol = np.where(mask, 100.0, 0.0).astype(np.float64)
    m = 25.0 * min(X, Y, Z) / 182.0
    for _ in range(nblobs):
        cx = rng.uniform(m, X - m)
        cy = rng.uniform(m, Y - m)
        cz = rng.uniform(m, Z - m)
        sigma = rng.uniform(1.5, 7.0)
        a = 60.0 * rng.uniform(-1.0, 1.0)
        _add_blob(vol, cx, cy, cz, sigma, a, mask)
    return np.ascontiguousarray(vol.astype(np.float32))


_NIFTI_CODES = {"uint8": (2, 8), "int16": (4, 16), "int32": (8, 32), "float32": (16, 32), "float64": (64, 64),
                "int8": (256, 8), "uint16": (512, 16), "uint32": (768, 32)}


def write_nifti(path, vol, pixdim=(1.0, 1.0, 1.0), qoffset=None, quatern=(0.0, 0.0, 0.0), dtype=np.float32):
    """Minimal single-file NIfTI-1 (.nii): 348-byte header + 4 pad bytes, vox_offset 352, voxels stored as
    ``dtype`` (float32 unless given).  With ``qoffset`` the qform (code 1) is written from ``quatern``
    (b, c, d) and the offsets."""
    import struct
    vol = np.ascontiguousarray(vol, dtype=dtype)
    code, bitpix = _NIFTI_CODES[vol.dtype.name]
    Z, Y, X = vol.shape
    h = bytearray(348)
    struct.pack_into("<i", h, 0, 348)
    struct.pack_into("<8h", h, 40, 3, X, Y, Z, 1, 1, 1, 1)
    struct.pack_into("<h", h, 70, code)    # datatype
    struct.pack_into("<h", h, 72, bitpix)  # bitpix
    struct.pack_into("<8f", h, 76, 1.0, pixdim[0], pixdim[1], pixdim[2], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", h, 108, 352.0)  # vox_offset
    struct.pack_into("<f", h, 112, 1.0)    # scl_slope
    if qoffset is not None:
        struct.pack_into("<h", h, 252, 1)  # qform_code
        struct.pack_into("<3f", h, 256, *quatern)
        struct.pack_into("<3f", h, 268, *qoffset)
    h[344:348] = b"n+1\0"
    with open(path, "wb") as f:
        f.write(bytes(h))
        f.write(b"\0\0\0\0")
        f.write(vol.tobytes())
