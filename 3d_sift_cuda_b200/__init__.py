"""B200-native 3D SIFT feature extraction (featExtract hot path of CarluerJB/3D_SIFT_CUDA).

The product is ``lib3dsift_b200.so`` (hand-written sm_100a kernels behind the C-ABI declared in
``include/s3d.h``).  This package is the thin Python host side used by the tests and the benchmark:
``api`` binds the C-ABI with ctypes (device memory and streams come from PyTorch), ``phantom``
generates the synthetic volumes of SURVEY.md section 8(d), ``featfile`` reads/writes the reference's
feature-file format, ``dist`` shards work across GPUs.

The directory name starts with a digit, so import it with
``importlib.import_module("3d_sift_cuda_b200")`` (see ``__graft_entry__.py``).
"""
from . import phantom  # noqa: F401
from .api import (  # noqa: F401
    CAND_DTYPE, FEATURE_DTYPE, KEYPOINT_DTYPE, Batch, Engine, Multi, Params, S3DError, build_library, gaussian_taps,
    library_path, load_library,
)
