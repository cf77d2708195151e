"""mlir-hashjoin_b200 — B200-native hash join behind the memref-in/memref-out surface of deveshv-99/mlir-HashJoin.

Layout (only what the build + probe path needs):
  csrc/        hand-written sm_100a kernels (hj_kernels.cu) and the C-ABI boundary (hj_capi.cu, include/hashjoin_b200.h)
  lib/         libhashjoin_b200.so, built in-tree by build.py (git-ignored, travels to the GPU box)
  _lib.py      ctypes binding of the C ABI; fails loudly when the library is missing — there is no CPU fallback
  join.py      host-side mirror of the reference's MLIR wrappers (@allocateHashTable ... @probeRelation, @main)
  datagen.py   seeded relation generators (device) for the BASELINE.json configs
  dist.py      multi-GPU plans: broadcast build and radix-partition + all-to-all (torch.distributed / NCCL)
"""
from .build import build_library, library_path  # noqa: F401

__all__ = ["build_library", "library_path"]
__version__ = "0.1"
