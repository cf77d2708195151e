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


# ---- the operators either side of the path --------------------------------------------------------------
def semi_join(probeRelation_: torch.Tensor, table: HashTable, probePayload: torch.Tensor | None = None, probeRowBase: int = 0) -> torch.Tensor:
    """Key-only mode: the row ids of the probe rows with at least one match in ``table``, each once (hjSemiJoinCount + hjSemiJoinWrite)."""
    _require_cuda(probeRelation_, "probeRelation")
    if not table.built:
        raise _lib.HashJoinError("semi_join on a table that was never built")
    lib = _lib.load()
    scratch = _scratch_for(table, probeRelation_)
    n = _lib.check_status(lib.hjSemiJoinCount(_ptr(probeRelation_), probeRelation_.numel(), table.key_bytes, _ptr(table.storage), _ptr(scratch), scratch.numel(),
                                              _ptr(probePayload), probeRowBase & 0xFFFFFFFF, _stream_ptr()), "hjSemiJoinCount")
    outS = torch.empty(n, dtype=torch.int32, device=probeRelation_.device)
    if n:
        _lib.check_status(lib.hjSemiJoinWrite(_ptr(probeRelation_), probeRelation_.numel(), table.key_bytes, _ptr(table.storage), _ptr(scratch), _ptr(outS),
                                              _ptr(probePayload), probeRowBase & 0xFFFFFFFF, _stream_ptr()), "hjSemiJoinWrite")
    return outS


def gather(column: torch.Tensor, rowIds: torch.Tensor, rowBase: int = 0) -> torch.Tensor:
    """Late materialisation of one payload column: out[k] = column[rowIds[k] - rowBase] (hjGather; 4- or 8-byte elements)."""
    _require_cuda(column, "column"); _require_cuda(rowIds, "rowIds")
    if rowIds.dtype != torch.int32 or column.element_size() not in (4, 8):
        raise _lib.HashJoinError("gather takes int32 row ids and a column of 4- or 8-byte elements")
    out = torch.empty(rowIds.numel(), dtype=column.dtype, device=column.device)
    rc = _lib.load().hjGather(_ptr(column), column.element_size(), _ptr(rowIds), rowIds.numel(), rowBase & 0xFFFFFFFF, _ptr(out), _stream_ptr())
    _lib.check_status(rc, "hjGather")
    return out


def nested_loop_join(table1: torch.Tensor, table2: torch.Tensor) -> torch.Tensor:
    """What nested-loop.mlir:203-283 computes — the equi-join of two row-major int32 tables on column 0, result rows = all columns of
    the larger table's row followed by the other table's non-key columns (:165-187) — as hash join + row materialisation on the GPU
    (hjExtractColumn, hjBuild / hjCount / hjWrite, hjMaterializeRows). The larger table is the outer (x) side like the reference (:252-262)."""
    for t in (table1, table2):
        _require_cuda(t, "table")
        if t.dtype != torch.int32 or t.dim() != 2:
            raise _lib.HashJoinError("nested_loop_join takes 2-D int32 tables (memref<?x?xi32>)")
    lib = _lib.load()
    x, y = (table2, table1) if table1.shape[0] < table2.shape[0] else (table1, table2)      # :252: the smaller table is the inner loop
    kx = torch.empty(x.shape[0], dtype=torch.int32, device=x.device); ky = torch.empty(y.shape[0], dtype=torch.int32, device=y.device)
    _lib.check_status(lib.hjExtractColumn(_ptr(x), x.shape[0], x.shape[1], 0, _ptr(kx), _stream_ptr()), "hjExtractColumn")
    _lib.check_status(lib.hjExtractColumn(_ptr(y), y.shape[0], y.shape[1], 0, _ptr(ky), _stream_ptr()), "hjExtractColumn")
    py, px = hash_join(ky, kx)                                                               # build on the smaller table y, probe with x
    out = torch.empty((px.numel(), x.shape[1] + y.shape[1] - 1), dtype=torch.int32, device=x.device)
    rc = lib.hjMaterializeRows(_ptr(x), x.shape[1], _ptr(y), y.shape[1], _ptr(px), _ptr(py), px.numel(), _ptr(out), _stream_ptr())
    _lib.check_status(rc, "hjMaterializeRows")
    return out


def pack_keys(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Two int32 key columns -> one int64 key column (hjPackKeys2x32): a two-column equi-join becomes an i64-key join."""
    _require_cuda(a, "a"); _require_cuda(b, "b")
    if a.dtype != torch.int32 or b.dtype != torch.int32 or a.numel() != b.numel():
        raise _lib.HashJoinError("pack_keys takes two int32 columns of equal length")
    out = torch.empty(a.numel(), dtype=torch.int64, device=a.device)
    _lib.check_status(_lib.load().hjPackKeys2x32(_ptr(a), _ptr(b), a.numel(), _ptr(out), _stream_ptr()), "hjPackKeys2x32")
    return out


def encode_float_keys(column: torch.Tensor, probe_side: bool) -> torch.Tensor:
    """f32 / f64 key column -> i32 / i64 keys with IEEE equality (hjEncodeFloatKeys): -0.0 joins +0.0, NaN joins nothing."""
    _require_cuda(column, "column")
    if column.dtype not in (torch.float32, torch.float64):
        raise _lib.HashJoinError("encode_float_keys takes a float32 or float64 column")
    eb = column.element_size()
    out = torch.empty(column.numel(), dtype=torch.int32 if eb == 4 else torch.int64, device=column.device)
    _lib.check_status(_lib.load().hjEncodeFloatKeys(_ptr(column), eb, column.numel(), 1 if probe_side else 0, _ptr(out), _stream_ptr()), "hjEncodeFloatKeys")
    return out


_SELECT_DTYPES = {torch.int32: 0, torch.int64: 1, torch.float32: 2, torch.float64: 3}
SELECT_OPS = {"<": 0, "<=": 1, ">": 2, ">=": 3, "==": 4, "!=": 5}


def selection(column: torch.Tensor, op: str, constant, rowBase: int = 0):
    """Experiments/selection.mlir:34-155 (`d_arrayA[i] < 80.0` there): (values, row ids) of the rows with ``value op constant``, in input
    order; count -> scan -> write through hjSelectCount / hjSelectWrite."""
    _require_cuda(column, "column")
    lib = _lib.load()
    dt = _SELECT_DTYPES[column.dtype]
    ic, fc = (int(constant), 0.0) if dt < 2 else (0, float(constant))
    n = column.numel()
    scratch = torch.empty(lib.hjSelectScratchBytes(n), dtype=torch.uint8, device=column.device)
    cnt = _lib.check_status(lib.hjSelectCount(_ptr(column), n, dt, SELECT_OPS[op], ic, fc, _ptr(scratch), scratch.numel(), _stream_ptr()), "hjSelectCount")
    vals = torch.empty(cnt, dtype=column.dtype, device=column.device); rows = torch.empty(cnt, dtype=torch.int32, device=column.device)
    _lib.check_status(lib.hjSelectWrite(_ptr(column), n, dt, SELECT_OPS[op], ic, fc, _ptr(scratch), _ptr(vals), _ptr(rows), rowBase & 0xFFFFFFFF, _stream_ptr()), "hjSelectWrite")
    return vals, rows


def pair_digest(outR: torch.Tensor, outS: torch.Tensor) -> tuple[int, int]:
    """Order-independent (sum, xor) digest of a device pair stream (K6)."""
    out = (C.c_uint64 * 2)()
    rc = _lib.load().hjPairDigest(_ptr(outR), _ptr(outS), outR.numel(), C.addressof(out), _stream_ptr())
    _lib.check_status(rc, "hjPairDigest")
    return int(out[0]), int(out[1])
