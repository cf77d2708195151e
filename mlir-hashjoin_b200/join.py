"""Host-side mirror of the reference's MLIR host wrappers, calling the CUDA path through the C ABI.

Names and argument meaning follow join_v1.mlir:
  allocateHashTable   :25-39    -> one opaque device workspace sized by hjTableBytes (was four gpu.allocs)
  initializeHashTable :54-75    -> folded into buildTable (K0 clear); kept as a no-op for call-sequence parity
  buildTable          :77-108   -> hjBuild
  countRows           :110-147  -> hjCount (returns the result size: the one mandatory host sync, :140-146)
  probeRelation       :149-176  -> hjWrite
  main                :525-649  -> H2D, build, count, allocate result, probe, D2H, check
torch is used for device memory and streams only; every kernel is in libhashjoin_b200.so. No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import torch

from . import _lib

_KEY_DTYPES = {torch.int32: 4, torch.int64: 8}


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _lib.HashJoinError(f"{name} must be a CUDA tensor: the join runs only on the GPU (no CPU fallback)")
    if not t.is_contiguous():
        raise _lib.HashJoinError(f"{name} must be contiguous (memref stride 1, shared.cpp:35)")


@dataclass
class HashTable:
    """Opaque B200 table workspace (replaces linkedListKey/RowId/nextIndex + hashTablePointers, join_v1.mlir:25-39)."""
    storage: torch.Tensor            # uint8, device
    key_bytes: int
    num_tuples: int
    built: bool = False
    scratch: torch.Tensor | None = field(default=None, repr=False)


def allocateHashTable(numTuples: int, hashTableSize: int | None = None, key_dtype: torch.dtype = torch.int32,
                      device: torch.device | str = "cuda") -> HashTable:
    """join_v1.mlir:25-39. ``hashTableSize`` (the reference's bucket count H) is accepted and ignored: the
    open-addressing table sizes itself from ``numTuples`` (load factor <= 0.5)."""
    lib = _lib.load()
    kb = _KEY_DTYPES[key_dtype]
    nbytes = _lib.check_status(lib.hjTableBytes(numTuples, kb), "hjTableBytes")
    return HashTable(torch.empty(nbytes, dtype=torch.uint8, device=device), kb, numTuples)


def initializeHashTable(table: HashTable) -> None:
    """join_v1.mlir:54-75. The clear (head[i] = -1 there, every slot EMPTY here) is the first launch of buildTable."""
    table.built = False


def buildTable(buildRelation: torch.Tensor, table: HashTable, payload: torch.Tensor | None = None, rowBase: int = 0) -> None:
    """join_v1.mlir:77-108. Row ids are ``rowBase + i`` (the reference stores the thread index, :232) unless a u32/i32
    ``payload`` column carries them (used by the radix-partitioned multi-GPU plan)."""
    _require_cuda(buildRelation, "buildRelation")
    if _KEY_DTYPES.get(buildRelation.dtype) != table.key_bytes:
        raise _lib.HashJoinError("build key dtype does not match the table")
    if payload is not None:
        _require_cuda(payload, "payload")
        if payload.dtype != torch.int32 or payload.numel() != buildRelation.numel():
            raise _lib.HashJoinError("payload must be an int32 column as long as the build relation")
    lib = _lib.load()
    rc = lib.hjBuild(_ptr(buildRelation), buildRelation.numel(), table.key_bytes, _ptr(payload), rowBase & 0xFFFFFFFF,
                     _ptr(table.storage), table.storage.numel(), _stream_ptr())
    _lib.check_status(rc, "hjBuild")
    table.built = True


def _scratch_for(table: HashTable, probeRelation: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    need = _lib.check_status(lib.hjScratchBytes(probeRelation.numel(), table.key_bytes), "hjScratchBytes")
    if table.scratch is None or table.scratch.numel() < need or table.scratch.device != probeRelation.device:
        table.scratch = torch.empty(need, dtype=torch.uint8, device=probeRelation.device)
    return table.scratch


def countRows(probeRelation: torch.Tensor, table: HashTable, probePayload: torch.Tensor | None = None, probeRowBase: int | None = None) -> int:
    """join_v1.mlir:110-147 -> result size. Runs K2 (count) + K3 (scan) and reads the total back (one host sync).
    ``probePayload`` / ``probeRowBase``: the probe row ids, when already known (hjCountRows): a slice-ordered copy of the probe
    relation then carries them; probeRelation() must be given the same ones."""
    _require_cuda(probeRelation, "probeRelation")
    if not table.built:
        raise _lib.HashJoinError("countRows on a table that was never built")
    if _KEY_DTYPES.get(probeRelation.dtype) != table.key_bytes:
        raise _lib.HashJoinError("probe key dtype does not match the table")
    lib = _lib.load()
    scratch = _scratch_for(table, probeRelation)
    if probePayload is not None or probeRowBase is not None:
        total = lib.hjCountRows(_ptr(probeRelation), probeRelation.numel(), table.key_bytes, _ptr(table.storage), _ptr(scratch), scratch.numel(),
                                _ptr(probePayload), (probeRowBase or 0) & 0xFFFFFFFF, _stream_ptr())
        return _lib.check_status(total, "hjCountRows")
    total = lib.hjCount(_ptr(probeRelation), probeRelation.numel(), table.key_bytes, _ptr(table.storage),
                        _ptr(scratch), scratch.numel(), _stream_ptr())
    return _lib.check_status(total, "hjCount")


def probeRelation(probeRelation_: torch.Tensor, table: HashTable, resultIndicesR: torch.Tensor, resultIndicesS: torch.Tensor,
                  probePayload: torch.Tensor | None = None, probeRowBase: int = 0) -> None:
    """join_v1.mlir:149-176. Fills the caller-allocated result columns (sized by countRows) with (build_row, probe_row)."""
    _require_cuda(probeRelation_, "probeRelation")
    for name, t in (("resultIndicesR", resultIndicesR), ("resultIndicesS", resultIndicesS)):
        _require_cuda(t, name)
        if t.dtype != torch.int32:
            raise _lib.HashJoinError(f"{name} must be int32 (join_v1.mlir:604-605)")
    if table.scratch is None:
        raise _lib.HashJoinError("probeRelation before countRows")
    lib = _lib.load()
    rc = lib.hjWrite(_ptr(probeRelation_), probeRelation_.numel(), table.key_bytes, _ptr(table.storage), _ptr(table.scratch),
                     _ptr(resultIndicesR), _ptr(resultIndicesS), _ptr(probePayload), probeRowBase & 0xFFFFFFFF, _stream_ptr())
    _lib.check_status(rc, "hjWrite")


def debug_sparse_flag(table: HashTable, probeRelation_: torch.Tensor) -> int:
    """Which probe path the last countRows took for this probe relation: 0 = match cache, 1 = hit lists (hjProbePath)."""
    lib = _lib.load()
    return _lib.check_status(lib.hjProbePath(_ptr(table.scratch), probeRelation_.numel(), table.key_bytes, _stream_ptr()), "hjProbePath")


def join_fused(probeRelation_: torch.Tensor, table: HashTable, resultIndicesR: torch.Tensor, resultIndicesS: torch.Tensor,
               probePayload: torch.Tensor | None = None, probeRowBase: int = 0) -> int:
    """Single-pass probe (hjJoinFused) into caller-allocated result columns whose length bounds the result (e.g. |S| for a
    unique build). Returns the number of pairs found; when it exceeds the columns' length only the first ``len`` were written.
    Raises for grouped tables (duplicate build keys): use countRows + probeRelation there."""
    _require_cuda(probeRelation_, "probeRelation")
    if not table.built:
        raise _lib.HashJoinError("join_fused on a table that was never built")
    lib = _lib.load()
    scratch = _scratch_for(table, probeRelation_)
    n = lib.hjJoinFused(_ptr(probeRelation_), probeRelation_.numel(), table.key_bytes, _ptr(table.storage), _ptr(scratch), scratch.numel(),
                        _ptr(resultIndicesR), _ptr(resultIndicesS), min(resultIndicesR.numel(), resultIndicesS.numel()),
                        _ptr(probePayload), probeRowBase & 0xFFFFFFFF, _stream_ptr())
    return _lib.check_status(n, "hjJoinFused")


def hash_join(buildRelation: torch.Tensor, probeRelation_: torch.Tensor, table: HashTable | None = None,
              buildPayload: torch.Tensor | None = None, probePayload: torch.Tensor | None = None,
              rowBase: int = 0, probeRowBase: int = 0):
    """Device-resident join: build -> count -> allocate -> write. Returns (resultIndicesR, resultIndicesS)."""
    if table is None:
        table = allocateHashTable(buildRelation.numel(), None, buildRelation.dtype, buildRelation.device)
    initializeHashTable(table)
    buildTable(buildRelation, table, buildPayload, rowBase)
    n = countRows(probeRelation_, table, probePayload, probeRowBase)                # the probe row ids are known here: hjCountRows
    outR = torch.empty(n, dtype=torch.int32, device=probeRelation_.device)          # join_v1.mlir:604-605
    outS = torch.empty(n, dtype=torch.int32, device=probeRelation_.device)
    if n != 0:                                                                       # :600-601
        probeRelation(probeRelation_, table, outR, outS, probePayload, probeRowBase)
    return outR, outS


def main(hostBuildRelation: torch.Tensor, hostProbeRelation: torch.Tensor, device: torch.device | str = "cuda", verify: bool = True):
    """join_v1.mlir:525-649 with caller-supplied host relations: H2D, join, D2H, check(). Returns
    (hostResultIndicesR, hostResultIndicesS, success) where success is check()'s tri-state (1/0/-1) or None."""
    dR = hostBuildRelation.to(device, non_blocking=True)                             # :558-561
    dS = hostProbeRelation.to(device, non_blocking=True)
    outR, outS = hash_join(dR, dS)
    hR, hS = outR.cpu(), outS.cpu()                                                  # :611-615
    success = None
    if verify and hostBuildRelation.dtype == torch.int32:
        success = check(hostBuildRelation, hostProbeRelation, hR, hS)               # :628
    return hR, hS, success


def check(hostBuildRelation: torch.Tensor, hostProbeRelation: torch.Tensor, hostResultIndicesR: torch.Tensor,
          hostResultIndicesS: torch.Tensor) -> int:
    """The library's check() (same tri-state verdicts as shared.cpp:129-172), through the expanded memref ABI."""
    lib = _lib.load()
    args = []
    for t in (hostBuildRelation, hostProbeRelation, hostResultIndicesR, hostResultIndicesS):
        if t.is_cuda or t.dtype != torch.int32 or not t.is_contiguous():
            raise _lib.HashJoinError("check() takes contiguous int32 HOST tensors")
        args += [t.data_ptr(), t.data_ptr(), 0, t.numel(), 1]
    return lib.check(*args)


def join_host(hostBuildRelation: torch.Tensor, hostProbeRelation: torch.Tensor, capacity: int | None = None):
    """End-to-end through hjJoinHost: host buffers in, host pairs out (H2D/D2H inside the call)."""
    lib = _lib.load()
    kb = _KEY_DTYPES[hostBuildRelation.dtype]
    n = lib.hjJoinHost(hostBuildRelation.data_ptr(), hostBuildRelation.numel(), hostProbeRelation.data_ptr(),
                       hostProbeRelation.numel(), kb, None, None, 0) if capacity is None else capacity
    n = _lib.check_status(n, "hjJoinHost")
    outR = torch.empty(n, dtype=torch.int32, pin_memory=torch.cuda.is_available())
    outS = torch.empty(n, dtype=torch.int32, pin_memory=torch.cuda.is_available())
    got = lib.hjJoinHost(hostBuildRelation.data_ptr(), hostBuildRelation.numel(), hostProbeRelation.data_ptr(),
                         hostProbeRelation.numel(), kb, outR.data_ptr(), outS.data_ptr(), n)
    got = _lib.check_status(got, "hjJoinHost")
    return outR[:got], outS[:got], got


def pair_digest(outR: torch.Tensor, outS: torch.Tensor) -> tuple[int, int]:
    """Order-independent (sum, xor) digest of a device pair stream (K6)."""
    out = (C.c_uint64 * 2)()
    rc = _lib.load().hjPairDigest(_ptr(outR), _ptr(outS), outR.numel(), C.addressof(out), _stream_ptr())
    _lib.check_status(rc, "hjPairDigest")
    return int(out[0]), int(out[1])
