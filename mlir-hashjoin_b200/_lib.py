"""ctypes binding of include/hashjoin_b200.h. The library is the product; if it is missing this module raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from .build import library_path

_i32, _i64, _u32, _u64, _vp = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_void_p
_MEMREF = [_vp, _vp, _i64, _i64, _i64]          # allocated, aligned, offset, size, stride  (shared.cpp:35)


class HjMemRef1D(C.Structure):                   # StridedMemRefType<T,1>
    _fields_ = [("allocated", _vp), ("aligned", _vp), ("offset", _i64), ("sizes", _i64 * 1), ("strides", _i64 * 1)]


_SIGNATURES = {
    # A. legacy helper symbols
    "startTimer": (None, []),
    "endTimer": (None, []),
    "initRelationIndex": (None, _MEMREF),
    "initRelationR": (None, _MEMREF),
    "initRelationS": (None, _MEMREF),
    "check": (_i32, _MEMREF * 4),
    "hashJoinSetSeeds": (None, [_u64, _u64]),
    # B. reference join entry points
    "initializeHashTable": (None, [_i64] + _MEMREF),
    "buildTable": (None, _MEMREF + [_i64] + _MEMREF * 4 + [_i32]),
    "countRows": (_i64, _MEMREF + [_i64] + _MEMREF * 5 + [_i32]),
    "probeRelation": (None, _MEMREF + [_i64, _i32] + _MEMREF * 7),
    "calculateNumberOfBlocks": (_i64, [_i64, _i64]),
    "hashJoinRelease": (None, []),
    "_mlir_ciface_initializeHashTable": (None, [_i64, _vp]),
    "_mlir_ciface_buildTable": (None, [_vp, _i64, _vp, _vp, _vp, _vp, _i32]),
    "_mlir_ciface_countRows": (_i64, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32]),
    "_mlir_ciface_probeRelation": (None, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "_mlir_ciface_check": (_i32, [_vp] * 4),
    "_mlir_ciface_initRelationIndex": (None, [_vp]),
    "_mlir_ciface_initRelationR": (None, [_vp]),
    "_mlir_ciface_initRelationS": (None, [_vp]),
    # C1. native MLIR surface
    "hashJoinTableBytes": (_i64, [_i64]),
    "hashJoinTableBytesI64": (_i64, [_i64]),
    "hashJoinScratchBytes": (_i64, [_i64]),
    "hashJoinScratchBytesI64": (_i64, [_i64]),
    "hashJoinBuild": (_i32, _MEMREF * 2),
    "hashJoinCount": (_i64, _MEMREF * 3),
    "hashJoinWrite": (_i32, _MEMREF * 5),
    "hashJoinBuildI64": (_i32, _MEMREF * 2),
    "hashJoinCountI64": (_i64, _MEMREF * 3),
    "hashJoinWriteI64": (_i32, _MEMREF * 5),
    "hashJoinGather": (_i32, _MEMREF * 3),
    "_mlir_ciface_hashJoinGather": (_i32, [_vp] * 3),
    "_mlir_ciface_hashJoinBuild": (_i32, [_vp] * 2),
    "_mlir_ciface_hashJoinCount": (_i64, [_vp] * 3),
    "_mlir_ciface_hashJoinWrite": (_i32, [_vp] * 5),
    "_mlir_ciface_hashJoinBuildI64": (_i32, [_vp] * 2),
    "_mlir_ciface_hashJoinCountI64": (_i64, [_vp] * 3),
    "_mlir_ciface_hashJoinWriteI64": (_i32, [_vp] * 5),
    # C2. native C surface
    "hjTableBytes": (_i64, [_i64, _i32]),
    "hjScratchBytes": (_i64, [_i64, _i32]),
    "hjBuild": (_i32, [_vp, _i64, _i32, _vp, _u32, _vp, _i64, _vp]),
    "hjBuildEx": (_i32, [_vp, _i64, _i32, _vp, _u32, _vp, _i64, _u32, _vp]),
    "hjDefaultPolicy": (_u32, []),
    "hjCountAsync": (_i32, [_vp, _i64, _i32, _vp, _vp, _i64, _vp]),
    "hjCountResult": (_i64, [_vp, _i64, _i32, _vp]),
    "hjCount": (_i64, [_vp, _i64, _i32, _vp, _vp, _i64, _vp]),
    "hjCountAsyncRows": (_i32, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _u32, _vp]),
    "hjCountRows": (_i64, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _u32, _vp]),
    "hjWrite": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _u32, _vp]),
    "hjJoinFused": (_i64, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _u32, _vp]),
    "hjPartitionWorkspaceBytes": (_i64, [_i64, _i32]),
    "hjPartition": (_i32, [_vp, _vp, _u32, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "hjPartitionCount": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp]),
    "hjPartitionPush": (_i32, [_vp, _vp, _u32, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "hjPairDigest": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "hjGenerate": (_i32, [_vp, _i64, _i32, _i32, _u64, _i64, _u64, _u32, _u64, _i64, _u64, _vp]),
    "hjSemiJoinCount": (_i64, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _u32, _vp]),
    "hjSemiJoinWrite": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _u32, _vp]),
    "hjGather": (_i32, [_vp, _i32, _vp, _i64, _u32, _vp, _vp]),
    "hjMaterializeRows": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _i64, _vp, _vp]),
    "hjExtractColumn": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "hjPackKeys2x32": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "hjEncodeFloatKeys": (_i32, [_vp, _i32, _i64, _i32, _vp, _vp]),
    "hjSelectScratchBytes": (_i64, [_i64]),
    "hjSelectCount": (_i64, [_vp, _i64, _i32, _i32, _i64, C.c_double, _vp, _i64, _vp]),
    "hjSelectWrite": (_i32, [_vp, _i64, _i32, _i32, _i64, C.c_double, _vp, _vp, _vp, _u32, _vp]),
    "hjSetHostChunkRows": (None, [_i64]),
    "hjGenerateAt": (_i32, [_vp, _vp, _i64, _i32, _i32, _u64, _i64, _u64, _u32, _u64, _u64, _vp]),
    "hjJoinHost": (_i64, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _i64]),
    "hjSetAllowDense": (None, [_i32]),
    "hjSetLocality": (None, [_i32]),
    "hjSetTmaCount": (None, [_i32]),
    "hjSetSparse": (None, [_i32]),
    "hjSetDenseWaves": (None, [_i32]),
    "hjSetDupSample": (None, [_i32]),
    "hjSetPartitionThreads": (None, [_i32]),
    "hjSetSliced": (None, [_i32]),
    "hjProbePath": (_i32, [_vp, _i64, _i32, _vp]),
    "hjTableLayout": (_i32, [_vp, _vp]),
    "hjLastErrorString": (C.c_char_p, []),
    "hjVersion": (C.c_char_p, []),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def load(path: str | Path | None = None) -> C.CDLL:
    """Load libhashjoin_b200.so and attach argtypes. Raises if the library has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else library_path()
    if not p.exists():
        raise RuntimeError(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The join has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(p))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = list(args)
    if path is None:
        _lib = lib
    return lib


class HashJoinError(RuntimeError):
    pass


def check_status(rc: int, what: str) -> int:
    if rc < 0:
        msg = load().hjLastErrorString()
        raise HashJoinError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")
    return rc
