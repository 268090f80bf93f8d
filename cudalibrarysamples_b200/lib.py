"""ctypes loader for libb200spmv.so (the product) and the real libcusparse.so.12 (handle management + GPU oracle).

The product path fails LOUDLY when the CUDA extension is missing: there is no CPU or library fallback here.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

REAL_CUSPARSE_CANDIDATES = (
    os.environ.get("B200SPMV_CUSPARSE", ""),
    "/usr/local/cuda/lib64/libcusparse.so.12",
    "libcusparse.so.12",
)

_shim = None
_real = None
_real_path = None


def real_cusparse_path() -> str:
    global _real_path
    if _real_path is None:
        for p in REAL_CUSPARSE_CANDIDATES:
            if not p:
                continue
            if "/" in p and not os.path.exists(p):
                continue
            _real_path = p
            break
        else:
            raise RuntimeError("real libcusparse.so.12 not found (set B200SPMV_CUSPARSE)")
    return _real_path


def real() -> C.CDLL:
    """The closed libcusparse.so.12: cusparseCreate/Destroy/SetStream/SetPointerMode, and the GPU oracle."""
    global _real
    if _real is None:
        _real = C.CDLL(real_cusparse_path(), mode=C.RTLD_LOCAL)
    return _real


def shim(build_if_missing: bool = False) -> C.CDLL:
    """libb200spmv.so. Raises if it has not been built (run __graft_entry__.build())."""
    global _shim
    if _shim is None:
        # the shim must dlopen the very same libcusparse instance that creates the handles we pass to it
        os.environ["B200SPMV_CUSPARSE"] = real_cusparse_path()
        path = _build.LIB_PATH
        if not os.path.exists(path):
            if build_if_missing:
                _build.build_native()
            else:
                raise RuntimeError(
                    f"{path} is missing: the CUDA extension was not built. Run `python -c 'import __graft_entry__ as g; "
                    "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        lib.b200spmv_csr_workspace_bytes.restype = C.c_size_t
        lib.b200spmv_coo_workspace_bytes.restype = C.c_size_t
        lib.b200spmv_sell_workspace_bytes.restype = C.c_size_t
        lib.b200spmv_csr_plan_tiles_offset.restype = C.c_size_t
        lib.b200spmv_csr_num_tiles.restype = C.c_int64
        lib.b200spmv_version.restype = C.c_char_p
        _shim = lib
    return _shim
