"""cudalibrarysamples_b200 -- B200-native (sm_100a) drop-in for the cusparseSpMV path of NVIDIA/CUDALibrarySamples.

  csrc/            hand-written CUDA kernels (CSR / COO / Sliced-ELL SpMV, tile partition) + the cuSPARSE-symbol shim
  cusparse_api.py  host-side mirror of the cuSPARSE generic API the samples call (ctypes over the C ABI)
  workloads.py     BASELINE.json's synthetic workloads, generated on the device
  sharded.py       row-block sharding of A over the GPUs of one box + allgather of x (torch.distributed / NCCL)
  build.py         nvcc build of libb200spmv.so (in-tree)
"""
from . import build  # noqa: F401

__all__ = ["build"]
__version__ = "0.1.0"
