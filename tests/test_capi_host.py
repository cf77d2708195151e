"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/hashjoin_b200.h declares, its host
helpers (check, initRelation*, timers) behave like shared_stuff/shared.cpp, and the join refuses to run without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _memref(a):
    p = a.ctypes.data_as(C.c_void_p)
    return [p, p, 0, a.size, 1]


def test_header_symbols_exported(lib):
    from mlir_hashjoin_b200 import _lib
    header = (ROOT / "include" / "hashjoin_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"^\s*(?:const\s+char\*|void|int32_t|int64_t|uint32_t)\s+(\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 50
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_check_matches_reference_golden(lib, golden):
    for c in golden:
        arrs = [np.array(c[k], np.int32) for k in ("R", "S", "outR", "outS")]
        args = sum((_memref(a) for a in arrs), [])
        assert lib.check(*args) == c["verdict"], c["name"]


def test_check_ciface_and_live_reference(lib, ref_check):
    from mlir_hashjoin_b200._lib import HjMemRef1D
    rng = np.random.default_rng(11)
    for k in range(40):
        R = rng.integers(-3, 4, int(rng.integers(0, 40))).astype(np.int32)
        S = rng.integers(-3, 4, int(rng.integers(0, 60))).astype(np.int32)
        i, j = np.nonzero(R[:, None] == S[None, :])
        a, b = i.astype(np.int32), j.astype(np.int32)
        if k % 3 == 1 and a.size:
            a = a[1:]; b = b[1:]
        if k % 3 == 2 and a.size:
            b = b.copy(); b[-1] ^= 1
        descs = []
        for arr in (R, S, a, b):
            d = HjMemRef1D(arr.ctypes.data, arr.ctypes.data, 0, (C.c_int64 * 1)(arr.size), (C.c_int64 * 1)(1))
            descs.append(d)
        got = lib._mlir_ciface_check(*[C.addressof(d) for d in descs])
        flat = lib.check(*sum((_memref(x) for x in (R, S, a, b)), []))
        assert got == flat
        if ref_check is not None:
            assert got == ref_check(R, S, a, b), k


def test_init_relations_and_blocks(lib):
    a = np.full(100, -7, np.int32)
    lib.initRelationIndex(*_memref(a))
    assert np.array_equal(a, np.arange(100))                                    # shared.cpp:35-41
    lib.hashJoinSetSeeds(5, 6)
    r1, r2, s1 = np.zeros(1000, np.int32), np.zeros(1000, np.int32), np.zeros(1000, np.int32)
    lib.initRelationR(*_memref(r1)); lib.initRelationR(*_memref(r2)); lib.initRelationS(*_memref(s1))
    assert np.array_equal(r1, r2) and not np.array_equal(r1, s1)
    assert r1.min() >= 1 and r1.max() <= 1_000_000_000                          # shared.cpp:13-14
    assert lib.calculateNumberOfBlocks(1000, 256) == 4 and lib.calculateNumberOfBlocks(1024, 256) == 4   # join_v1.mlir:43-52
    assert lib.calculateNumberOfBlocks(0, 256) == 0


def test_timer_format(lib, capfd):
    lib.startTimer(); lib.endTimer()
    out = capfd.readouterr().out
    assert re.search(r"For \d+, time taken: \d+ microseconds", out)             # shared.cpp:28-29


def test_workspace_queries(lib):
    assert lib.hjTableBytes(0, 4) >= 256 + 64
    # room for whichever layout the build picks: inline buckets at load 0.5 (16 / 32 bytes per row), the grouped layout for duplicate
    # keys (u32 row ids + 16-byte slots at load <= 0.8: 24 bytes per row) or, beyond L2 reach, the radix layout (two copies of
    # (key, row id) + offsets + partition workspace)
    for n in (1 << 10, 1 << 20):
        assert 24 * n <= lib.hjTableBytes(n, 4) - 256 <= 24 * n + 1024
        assert 32 * n <= lib.hjTableBytes(n, 8) - 256 <= 32 * n + 1024
    assert lib.hjTableBytes(1 << 24, 4) - 256 >= 24 * (1 << 24)
    assert lib.hjTableBytes(1 << 24, 8) - 256 >= 2 * 12 * (1 << 24)
    assert lib.hjScratchBytes(1 << 24, 8) >= (8 + 2 * 12) * (1 << 24)
    assert lib.hjTableBytes(10, 5) < 0 and lib.hjScratchBytes(-1, 4) < 0
    assert lib.hashJoinTableBytes(1000) == lib.hjTableBytes(1000, 4)
    assert lib.hashJoinScratchBytesI64(1000) == lib.hjScratchBytes(1000, 8)
    assert lib.hjScratchBytes(1 << 20, 4) >= 4 << 20


def test_no_cpu_fallback(lib):
    """Without a CUDA device every join entry point must fail loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    R = np.arange(16, dtype=np.int32); S = np.arange(16, dtype=np.int32)
    rc = lib.hjJoinHost(R.ctypes.data, 16, S.ctypes.data, 16, 4, None, None, 0)
    assert rc < 0
    assert lib.hjLastErrorString()
    from mlir_hashjoin_b200 import _lib, join
    with pytest.raises(_lib.HashJoinError):
        join.buildTable(torch.arange(4, dtype=torch.int32), join.HashTable(torch.empty(1024, dtype=torch.uint8), 4, 4))


def test_argument_validation_of_native_entry_points(lib):
    """Bad arguments are refused on the host, before any CUDA call: negative status, a message in hjLastErrorString, nothing thrown."""
    buf = np.zeros(4096, np.uint8)
    p = buf.ctypes.data
    assert lib.hjScratchBytes(1 << 20, 4) >= 8 << 20                     # match cache + as much again for the hit lists / run starts
    assert lib.hjCountRows(None, 10, 4, p, p, buf.size, None, 0, None) < 0          # null probe relation
    assert lib.hjCountRows(p, 10, 3, p, p, buf.size, None, 0, None) < 0             # key width
    assert lib.hjCountAsyncRows(p, (1 << 32) + 1, 4, p, p, buf.size, None, 0, None) < 0   # more probe rows than 32-bit row ids
    assert lib.hjCountRows(p, 1 << 20, 4, p, p, 1024, None, 0, None) < 0            # scratch too small
    assert lib.hjJoinFused(p, 10, 4, p, None, 0, p, p, 10, None, 0, None) < 0        # no scratch
    assert lib.hjJoinFused(p, 10, 4, p, p, buf.size, None, None, 10, None, 0, None) < 0   # capacity without result columns
    assert lib.hjTableLayout(None, None) < 0
    # workspace alignment: buckets are 32-byte vector loads in 64-byte pairs; a table at +16 must be refused, not faulted on
    base = (p + 255) & ~255
    assert lib.hjBuild(base, 16, 4, None, 0, base + 16, 2048, None) == -22 and b"64-byte" in lib.hjLastErrorString()
    assert lib.hjBuildEx(base, 16, 4, None, 0, base + 1024, 2048, 3, None) == -22          # unknown policy bits (before any CUDA call)
    assert lib.hjCountAsync(base, 16, 4, base + 16, base + 1024, 1 << 30, None) == -22
    assert lib.hjCountAsync(base, 16, 4, base + 1024, base + 16, 1 << 30, None) == -22
    assert lib.hjWrite(base, 16, 4, base + 16, base + 1024, base, base, None, 0, None) == -22
    # row ids are 32-bit and 0xFFFFFFFF is the EMPTY marker: a row base that would reach it is refused
    assert lib.hjBuild(base, 16, 4, None, 0xFFFFFFF0, base + 1024, 2048, None) == -22
    assert lib.hjJoinHost(p, 16, p, 1 << 32, 4, None, None, 0) == -22                      # probe rows beyond 32-bit row ids
    assert lib.hjEncodeFloatKeys(p, 2, 10, 0, p, None) == -22 and lib.hjEncodeFloatKeys(None, 4, 10, 0, p, None) == -22   # f32 / f64 only; null column
    assert lib.hjEncodeFloatKeys(None, 8, 0, 1, None, None) == 0                           # nothing to encode: no CUDA call
    assert lib.hjDefaultPolicy() & 3 in (0, 1, 2)
    assert lib.hjProbePath(None, 10, 4, None) < 0
    assert lib.hjPartitionWorkspaceBytes(1 << 20, 8) > 0
    assert lib.hjLastErrorString()
    for setter in (lib.hjSetAllowDense, lib.hjSetSparse, lib.hjSetDupSample):  # switches are plain host state: callable without a device
        setter(1)
    lib.hjSetAllowDense(2)
