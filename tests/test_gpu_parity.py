"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.
Bit-exact bar: identical multiset of (build_row, probe_row) i32 pairs after a canonical sort (shared.cpp:168-171)."""
import ctypes as C

import numpy as np
import pytest

from oracle.binding import sorted_pairs

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["auto", "hash", "cache", "lists", "hash-lists"])
def layout(request, lib):
    """Every parity case runs under each table / probe policy: the default (direct-address layout for dense key ranges, counted by
    range test when the range is gap-free and unique, falling back to the hash layout on a duplicate), the bucketised hash layout
    forced, the direct-address layout with the match cache only, and the last two again with the hit-list probe path
    (hjSetSparse(2)) forced instead of the match cache."""
    lib.hjSetAllowDense({"auto": 2, "hash": 0, "cache": 1, "lists": 1, "hash-lists": 0}[request.param])
    lib.hjSetSparse(2 if "lists" in request.param else 1)
    lib.hjSetDupSample(0 if request.param == "hash" else 1)       # "hash" keeps the attempt-inline-then-abort path under test
    yield request.param
    lib.hjSetAllowDense(2)
    lib.hjSetSparse(1)
    lib.hjSetDupSample(1)


def _join_np(R, S, cuda):
    import torch
    from mlir_hashjoin_b200 import join
    dR = torch.from_numpy(np.ascontiguousarray(R)).to(cuda)
    dS = torch.from_numpy(np.ascontiguousarray(S)).to(cuda)
    a, b = join.hash_join(dR, dS)
    torch.cuda.synchronize()
    return a.cpu().numpy(), b.cpu().numpy()


def _assert_parity(oracle, R, S, a, b, H=None):
    oa, ob = oracle.join(R, S, H=H)
    assert a.size == oa.size, f"count differs: gpu {a.size} oracle {oa.size}"
    assert np.array_equal(sorted_pairs(a, b), sorted_pairs(oa, ob))


def test_golden_fixtures(lib, cuda, oracle, golden, ref_check):
    """Every fixture the reference's check() accepted: the GPU join reproduces the accepted multiset exactly."""
    n = 0
    for c in golden:
        if c["verdict"] != 1 or "padded" in c["name"] or "nonempty" in c["name"]:
            continue
        R, S = np.array(c["R"], np.int32), np.array(c["S"], np.int32)
        a, b = _join_np(R, S, cuda)
        assert np.array_equal(sorted_pairs(a, b), sorted_pairs(c["outR"], c["outS"])), c["name"]
        assert oracle.check(R, S, a, b) == 1
        if ref_check is not None:
            assert ref_check(R, S, a, b) == 1, c["name"]
        n += 1
    assert n >= 15


def test_c1_config(lib, cuda, oracle, ref_check):
    """BASELINE.json config 1: 1K x 4K, i32 unique build keys; GPU generators are bit-identical to the oracle's."""
    import torch
    from mlir_hashjoin_b200 import datagen, join
    cfg = datagen.config("C1")
    dR, dS = datagen.generate(cfg.build), datagen.generate(cfg.probe)
    R = oracle.generate(1024, 4, 1, 1, 0, 1024); S = oracle.generate(4096, 4, 2, 2, 0, 2048)
    assert np.array_equal(dR.cpu().numpy(), R) and np.array_equal(dS.cpu().numpy(), S)
    a, b = join.hash_join(dR, dS)
    a, b = a.cpu().numpy(), b.cpu().numpy()
    _assert_parity(oracle, R, S, a, b, H=5)
    assert oracle.check(R, S, a, b) == 1
    if ref_check is not None:
        assert ref_check(R, S, a, b) == 1
    assert join.check(dR.cpu(), dS.cpu(), torch.from_numpy(a), torch.from_numpy(b)) == 1


@pytest.mark.parametrize("nR,nS,dom", [(1, 1, 1), (1, 5000, 3), (5000, 1, 3), (2047, 2049, 4000), (2048, 4096, 100), (100000, 300001, 150000),
                                       (3, 7, 2), (65, 6145, 64)])
def test_ragged_sizes_i32(lib, cuda, oracle, nR, nS, dom):
    rng = np.random.default_rng(nR * 31 + nS)
    R = rng.integers(-dom, dom, nR).astype(np.int32)            # duplicates on both sides, negative keys
    S = rng.integers(-dom, dom, nS).astype(np.int32)
    a, b = _join_np(R, S, cuda)
    _assert_parity(oracle, R, S, a, b)


def test_unique_build_partial_hits(lib, cuda, oracle):
    rng = np.random.default_rng(5)
    R = rng.permutation(200000).astype(np.int32)[:120000]
    S = rng.integers(0, 400000, 777777).astype(np.int32)
    a, b = _join_np(R, S, cuda)
    _assert_parity(oracle, R, S, a, b)


def test_selective_join_takes_hit_lists(lib, cuda, oracle, layout):
    """Config 3 in small: 2^16 unique build keys, 2^21 + 5 probe rows of which ~10 % hit. Under the default policy the device-side
    sample must pick the hit-list path (and the match cache when half of the rows hit); both give the oracle's multiset."""
    import torch
    from mlir_hashjoin_b200 import join
    rng = np.random.default_rng(33)
    nR, nS = 1 << 16, (1 << 21) + 5
    R = (rng.permutation(4 * nR)[:nR] * 3 + 7).astype(np.int32)
    for frac, expect_lists in ((0.1, True), (0.6, False)):
        hit = rng.random(nS) < frac
        S = np.where(hit, rng.choice(R, nS), rng.integers(1 << 28, 1 << 30, nS)).astype(np.int32)
        dR, dS = torch.from_numpy(R).to(cuda), torch.from_numpy(S).to(cuda)
        table = join.allocateHashTable(nR, None, dR.dtype, cuda)
        join.buildTable(dR, table)
        n = join.countRows(dS, table)
        sv_flag = join.debug_sparse_flag(table, dS)
        if layout in ("auto", "hash", "cache"):
            assert sv_flag == int(expect_lists), (frac, sv_flag)
        elif "lists" in layout:
            assert sv_flag == 1
        a = torch.empty(n, dtype=torch.int32, device=cuda); b = torch.empty(n, dtype=torch.int32, device=cuda)
        join.probeRelation(dS, table, a, b)
        _assert_parity(oracle, R, S, a.cpu().numpy(), b.cpu().numpy())


def test_tma_staged_count_kernel(lib, cuda, oracle):
    """hjSetTmaCount(1): the experimental direct-address count kernel with cp.async.bulk staged streams (off by default because it
    is slower) must give the same result, including the ragged last tile that falls back to LDG."""
    rng = np.random.default_rng(4)
    lib.hjSetTmaCount(1)
    try:
        for nR, nS in ((1024, 4096), (5000, 2048 * 8 * 3 + 777), (70000, 200003)):
            R = rng.permutation(2 * nR).astype(np.int32)[:nR]
            S = rng.integers(-5, 2 * nR + 5, nS).astype(np.int32)
            a, b = _join_np(R, S, cuda)
            _assert_parity(oracle, R, S, a, b)
    finally:
        lib.hjSetTmaCount(0)


def test_row_id_limits(lib, cuda):
    """Row ids are 32-bit patterns (join_v1.mlir:604-605): relations with more rows than that are refused, not wrapped."""
    import torch
    t = torch.empty(1 << 12, dtype=torch.uint8, device=cuda)
    assert lib.hjBuild(t.data_ptr(), (1 << 32), 4, None, 0, t.data_ptr(), t.numel(), None) < 0
    assert lib.hjCountAsync(t.data_ptr(), (1 << 32) + 5, 4, t.data_ptr(), t.data_ptr(), t.numel(), None) < 0


def test_fused_single_pass(lib, cuda, oracle):
    """hjJoinFused: lookup + decoupled look-back + write in one pass gives the same multiset as count + write; a too-small result
    only truncates (the count is still exact); grouped tables are refused."""
    import torch
    from mlir_hashjoin_b200 import _lib, join
    rng = np.random.default_rng(17)
    for kd, nR, nS in ((np.int32, 50_000, 300_007), (np.int64, 20_000, 100_003), (np.int32, 1, 5000), (np.int32, 300_000, 2_000_003)):
        R = rng.permutation(3 * nR)[:nR].astype(kd)                       # unique; sparse enough that the hash layout runs under "hash"
        if kd == np.int64:
            R = R * np.int64(0x9E3779B97F4A7C15 - (1 << 64))
        S = rng.choice(np.concatenate([R, R + 1]), nS).astype(kd)        # about half of the probe rows hit
        dR, dS = torch.from_numpy(R).to(cuda), torch.from_numpy(S).to(cuda)
        table = join.allocateHashTable(nR, None, dR.dtype, cuda)
        join.buildTable(dR, table)
        outR = torch.empty(nS, dtype=torch.int32, device=cuda); outS = torch.empty(nS, dtype=torch.int32, device=cuda)
        n = join.join_fused(dS, table, outR, outS, probeRowBase=11)
        oa, ob = oracle.join(R, S)
        assert n == oa.size
        assert np.array_equal(sorted_pairs(outR[:n].cpu().numpy(), outS[:n].cpu().numpy()), sorted_pairs(oa, ob + 11))
        small = torch.empty(max(1, n // 3), dtype=torch.int32, device=cuda)
        assert join.join_fused(dS, table, small, small.clone()) == n      # capacity too small: exact count, truncated output
    Rd = np.repeat(np.arange(1000, dtype=np.int32), 3)
    table = join.allocateHashTable(Rd.size, None, torch.int32, cuda)
    join.buildTable(torch.from_numpy(Rd).to(cuda), table)
    with pytest.raises(_lib.HashJoinError):
        join.join_fused(torch.from_numpy(Rd).to(cuda), table, outR, outS)


def test_empty_inputs(lib, cuda, oracle):
    e = np.empty(0, np.int32)
    for R, S in ((e, np.arange(10, dtype=np.int32)), (np.arange(10, dtype=np.int32), e), (e, e)):
        a, b = _join_np(R, S, cuda)
        assert a.size == 0 and b.size == 0
    a, b = _join_np(np.full(12, 10, np.int32), np.full(12, 1, np.int32), cuda)       # KAT-none: resultSize == 0 (join_v1.mlir:635-644)
    assert a.size == 0


def test_extreme_keys_no_reserved_value(lib, cuda, oracle):
    """The reference reserves no key value (its -1 sentinel is a node index, join_v1.mlir:197,333): 0xFFFFFFFF etc. must join."""
    R = np.array([-1, -1, -2**31, 2**31 - 1, 0, -1], np.int32)
    S = np.array([-1, 0, 2**31 - 1, -2**31, 5, -1, -1], np.int32)
    a, b = _join_np(R, S, cuda)
    _assert_parity(oracle, R, S, a, b)
    R64 = np.array([-1, 2**63 - 1, -2**63, 0, -1, 1 << 32, 1], np.int64)
    S64 = np.array([1 << 32, -1, 1, 2**63 - 1, -2**63, 0x100000001], np.int64)
    a, b = _join_np(R64, S64, cuda)
    _assert_parity(oracle, R64, S64, a, b)


@pytest.mark.parametrize("nR,nS", [(1, 3), (1023, 1025), (50000, 200003)])
def test_i64_keys(lib, cuda, oracle, nR, nS):
    rng = np.random.default_rng(nR + nS)
    base = rng.integers(0, max(2, nR // 2), nR).astype(np.int64)
    R = base * np.int64(0x9E3779B97F4A7C15 - (1 << 64))           # spread over the full 64-bit space (wraps)
    S = rng.integers(0, max(2, nR), nS).astype(np.int64) * np.int64(0x9E3779B97F4A7C15 - (1 << 64))
    a, b = _join_np(R, S, cuda)
    _assert_parity(oracle, R, S, a, b)


def test_heavy_duplicates(lib, cuda, oracle):
    """join-performances.md shape in miniature: few distinct keys, many matches per probe row (output-heavy)."""
    rng = np.random.default_rng(9)
    R = rng.integers(1, 101, 20000).astype(np.int32)
    S = rng.integers(1, 101, 5000).astype(np.int32)
    a, b = _join_np(R, S, cuda)
    assert a.size > 900000
    oa, ob = oracle.join(R, S, H=100, threads=4)
    assert a.size == oa.size
    assert oracle.pair_digest(a, b) == oracle.pair_digest(oa, ob)
    assert np.array_equal(sorted_pairs(a, b), sorted_pairs(oa, ob))


def test_unaligned_relation_pointers(lib, cuda, oracle):
    """memref offset != 0 (aligned + offset is only 4-byte aligned): the scalar-load variants must give the same result."""
    import torch
    from mlir_hashjoin_b200 import join
    rng = np.random.default_rng(21)
    R = rng.integers(0, 3000, 5001).astype(np.int32); S = rng.integers(0, 3000, 9003).astype(np.int32)
    dR = torch.from_numpy(R).to(cuda)[1:]; dS = torch.from_numpy(S).to(cuda)[3:]
    assert dR.data_ptr() % 16 != 0 and dS.data_ptr() % 16 != 0
    a, b = join.hash_join(dR.contiguous() if not dR.is_contiguous() else dR, dS)
    _assert_parity(oracle, R[1:], S[3:], a.cpu().numpy(), b.cpu().numpy())


def test_payload_columns_and_row_base(lib, cuda, oracle):
    """Row ids carried as payload columns (the radix-partitioned plan) and probe row bases (the broadcast plan)."""
    import torch
    from mlir_hashjoin_b200 import join
    rng = np.random.default_rng(33)
    R = rng.integers(0, 500, 3000).astype(np.int32); S = rng.integers(0, 500, 7000).astype(np.int32)
    pr = rng.permutation(3000).astype(np.int32) + 100000; ps = rng.permutation(7000).astype(np.int32) + 500000
    dR, dS = torch.from_numpy(R).to(cuda), torch.from_numpy(S).to(cuda)
    a, b = join.hash_join(dR, dS, buildPayload=torch.from_numpy(pr).to(cuda), probePayload=torch.from_numpy(ps).to(cuda))
    oa, ob = oracle.join(R, S)
    assert np.array_equal(sorted_pairs(a.cpu().numpy(), b.cpu().numpy()), sorted_pairs(pr[oa], ps[ob]))
    a, b = join.hash_join(dR, dS, rowBase=1000, probeRowBase=2000000)
    assert np.array_equal(sorted_pairs(a.cpu().numpy(), b.cpu().numpy()), sorted_pairs(oa + 1000, ob + 2000000))


def _dev_memref(t):
    return [t.data_ptr(), t.data_ptr(), 0, t.numel(), 1]


def test_reference_entry_points_expanded_abi(lib, cuda, oracle, capfd):
    """The reference's own call sequence (join_v1.mlir:565-615) through the expanded memref ABI with the reference's
    argument lists: caller-allocated chained-table arrays are only a handle; the result passes check()."""
    import torch
    rng = np.random.default_rng(77)
    nR, nS, H = 6000, 20000, 1000
    R = rng.integers(1, 4000, nR).astype(np.int32); S = rng.integers(1, 4000, nS).astype(np.int32)
    dR, dS = torch.from_numpy(R).to(cuda), torch.from_numpy(S).to(cuda)
    lkey = torch.empty(nR, dtype=torch.int32, device=cuda); lrow = torch.empty(nR, dtype=torch.int64, device=cuda)
    lnext = torch.empty(nR, dtype=torch.int64, device=cuda); head = torch.empty(H, dtype=torch.int32, device=cuda)
    prefix = torch.empty(nS, dtype=torch.int64, device=cuda)
    table = _dev_memref(head) + _dev_memref(lkey) + _dev_memref(lrow) + _dev_memref(lnext)
    lib.initializeHashTable(H, *_dev_memref(head))
    assert int(head.min()) == -1 and int(head.max()) == -1                           # join_v1.mlir:197
    lib.buildTable(*_dev_memref(dR), nR, *table, H)
    n = lib.countRows(*_dev_memref(dS), nS, *table, *_dev_memref(prefix), H)
    oa, ob = oracle.join(R, S, H=H)
    assert n == oa.size
    outR = torch.empty(n, dtype=torch.int32, device=cuda); outS = torch.empty(n, dtype=torch.int32, device=cuda)
    lib.probeRelation(*_dev_memref(dS), nS, H, *table, *_dev_memref(prefix), *_dev_memref(outR), *_dev_memref(outS))
    a, b = outR.cpu().numpy(), outS.cpu().numpy()
    assert np.array_equal(sorted_pairs(a, b), sorted_pairs(oa, ob))
    args = []
    for arr in (R, S, a, b):
        p = arr.ctypes.data_as(C.c_void_p); args += [p, p, 0, arr.size, 1]
    assert lib.check(*args) == 1
    out = capfd.readouterr().out
    assert out.count("time taken:") >= 4                                             # the four timer prints (For 0..3)
    # countRows on a table that was never built fails loudly
    other = torch.empty(8, dtype=torch.int32, device=cuda)
    bad = _dev_memref(other) + _dev_memref(lkey) + _dev_memref(lrow) + _dev_memref(lnext)
    assert lib.countRows(*_dev_memref(dS), nS, *bad, *_dev_memref(prefix), H) < 0
    lib.hashJoinRelease()


def test_native_mlir_surface_ciface(lib, cuda, oracle):
    """hashJoinBuild/Count/Write through llvm.emit_c_interface descriptors, i32 and i64."""
    import torch
    from mlir_hashjoin_b200._lib import HjMemRef1D
    rng = np.random.default_rng(99)

    def desc(t, off=0):
        return HjMemRef1D(t.data_ptr(), t.data_ptr(), off, (C.c_int64 * 1)(t.numel() - off), (C.c_int64 * 1)(1))
    for kd, suf in ((np.int32, ""), (np.int64, "I64")):
        R = rng.integers(0, 900, 2500).astype(kd); S = rng.integers(0, 900, 8100).astype(kd)
        dR, dS = torch.from_numpy(R).to(cuda), torch.from_numpy(S).to(cuda)
        tb = getattr(lib, "hashJoinTableBytes" + suf)(R.size); sb = getattr(lib, "hashJoinScratchBytes" + suf)(S.size)
        table = torch.empty(tb, dtype=torch.uint8, device=cuda); scratch = torch.empty(sb, dtype=torch.uint8, device=cuda)
        dsc = [desc(dR), desc(dS), desc(table), desc(scratch)]
        pR, pS, pT, pSc = [C.addressof(d) for d in dsc]
        assert getattr(lib, "_mlir_ciface_hashJoinBuild" + suf)(pR, pT) == 0
        n = getattr(lib, "_mlir_ciface_hashJoinCount" + suf)(pS, pT, pSc)
        oa, ob = oracle.join(R, S)
        assert n == oa.size
        outR = torch.empty(n, dtype=torch.int32, device=cuda); outS = torch.empty(n, dtype=torch.int32, device=cuda)
        dR_, dS_ = desc(outR), desc(outS)
        assert getattr(lib, "_mlir_ciface_hashJoinWrite" + suf)(pS, pT, pSc, C.addressof(dR_), C.addressof(dS_)) == 0
        assert np.array_equal(sorted_pairs(outR.cpu().numpy(), outS.cpu().numpy()), sorted_pairs(oa, ob))
    # too-small workspaces are rejected, not overrun
    small = torch.empty(64, dtype=torch.uint8, device=cuda)
    ds = desc(small)
    assert lib._mlir_ciface_hashJoinBuild(pR, C.addressof(ds)) < 0


def test_join_host_e2e(lib, cuda, oracle):
    import torch
    from mlir_hashjoin_b200 import join
    rng = np.random.default_rng(13)
    R = rng.permutation(50000).astype(np.int32); S = rng.integers(0, 80000, 200000).astype(np.int32)
    a, b, n = join.join_host(torch.from_numpy(R), torch.from_numpy(S))
    _assert_parity(oracle, R, S, a.numpy(), b.numpy())
    hR, hS, ok = join.main(torch.from_numpy(R), torch.from_numpy(S))
    assert ok == 1 and hR.numel() == n
    # more than one 2^24-row probe chunk: the H2D / join / D2H pipeline appends chunk results at a running offset
    S2 = rng.integers(0, 60000, (1 << 24) + 100003).astype(np.int32)
    a, b, n2 = join.join_host(torch.from_numpy(R), torch.from_numpy(S2))
    oa, ob = oracle.join(R, S2, threads=0)
    assert n2 == oa.size and oracle.pair_digest(a.numpy(), b.numpy()) == oracle.pair_digest(oa, ob)
    assert lib.hjJoinHost(R.ctypes.data, R.size, S2.ctypes.data, S2.size, 4, a.data_ptr(), b.data_ptr(), 1000) == n2   # too small: size only


def test_generators_bit_identical(lib, cuda, oracle):
    from mlir_hashjoin_b200 import datagen
    specs = [datagen.RelationSpec(10007, 4, datagen.KIND_UNIQUE, 42, 5, 20000), datagen.RelationSpec(10007, 4, datagen.KIND_UNIFORM, 43, -9, 777),
             datagen.RelationSpec(10007, 4, datagen.KIND_MIXED, 45, 0, 4096, 6554), datagen.RelationSpec(8192, 8, datagen.KIND_FK, 46, 0, 2048, 0, datagen.ODD_MUL64),
             datagen.RelationSpec(10007, 8, datagen.KIND_ZIPF, 47, 0, 1 << 12, 0, datagen.ODD_MUL64), datagen.RelationSpec(4099, 4, datagen.KIND_INDEX, 0),
             datagen.RelationSpec(10007, 8, datagen.KIND_UNIQUE, 48, 0, 10007, 0, datagen.ODD_MUL64), datagen.RelationSpec(5000, 4, datagen.KIND_UNIQUE, 5, 0, 5000, 0, 0x9E3779B1)]
    for s in specs:
        g = datagen.generate(s).cpu().numpy()
        o = oracle.generate(s.n, s.key_bytes, s.kind, s.seed, s.lo, s.domain, s.p16, s.key_mul)
        assert np.array_equal(g, o), s
        half = datagen.generate(s, index_base=1000, n_local=2000).cpu().numpy()     # sharded generation = slice of the whole
        assert np.array_equal(half, o[1000:3000]), s


def test_pair_digest_matches_oracle(lib, cuda, oracle):
    import torch
    from mlir_hashjoin_b200 import join
    rng = np.random.default_rng(1)
    a = rng.integers(-2**31, 2**31, 100003).astype(np.int32); b = rng.integers(-2**31, 2**31, 100003).astype(np.int32)
    assert join.pair_digest(torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)) == oracle.pair_digest(a, b)


@pytest.fixture
def scatter_shape(lib, request):
    """CTA shape of the partition scatter kernel: 0 = chosen by fan, else 256 / 512 / 1024 threads."""
    lib.hjSetPartitionThreads(request.param)
    yield request.param
    lib.hjSetPartitionThreads(0)


@pytest.mark.parametrize("scatter_shape", [0, 256, 512, 1024], indirect=True)
def test_radix_partition(lib, cuda, oracle, scatter_shape):
    """K5: every key lands in exactly one partition, equal keys in the same one, (key,row) pairs preserved."""
    import torch
    rng = np.random.default_rng(8)
    for kd, kb, n, parts in ((np.int32, 4, 100003, 8), (np.int64, 8, 100003, 8), (np.int64, 8, 300_007, 2), (np.int32, 4, 70_001, 5), (np.int64, 8, 3_000_017, 3),
                             (np.int32, 4, 2_000_003, 200), (np.int64, 8, 1_500_001, 256), (np.int64, 8, 5, 3), (np.int32, 4, 8192, 16), (np.int64, 8, 16385, 16)):
        keys = rng.integers(0, max(5000, 400 * parts), n).astype(kd)
        d = torch.from_numpy(keys).to(cuda)
        ok = torch.empty_like(d); orow = torch.empty(n, dtype=torch.int32, device=cuda)
        offs = torch.empty(parts + 1, dtype=torch.int64, device=cuda)
        ws = torch.empty(lib.hjPartitionWorkspaceBytes(n, parts), dtype=torch.uint8, device=cuda)
        rc = lib.hjPartition(d.data_ptr(), None, 7, n, kb, parts, ok.data_ptr(), orow.data_ptr(), offs.data_ptr(), ws.data_ptr(), ws.numel(), None)
        assert rc == 0
        torch.cuda.synchronize()
        offs = offs.cpu().numpy(); pk = ok.cpu().numpy(); prow = orow.cpu().numpy()
        assert offs[0] == 0 and offs[-1] == n and np.all(np.diff(offs) >= 0)
        assert np.array_equal(keys[prow - 7], pk)                                  # rows carried with their keys
        assert np.array_equal(np.sort(prow - 7), np.arange(n))                      # a permutation of the input
        owner = {}
        for p in range(parts):
            for k in np.unique(pk[offs[p]:offs[p + 1]]):
                assert owner.setdefault(int(k), p) == p                             # equal keys -> same partition
        sizes = np.diff(offs)
        if n > 50 * parts:
            assert sizes.min() > 0.5 * n / parts


def test_mid_size_vs_oracle(lib, cuda, oracle):
    """1M x 16M unique build, 50 % hits: count and order-independent digest against the OpenMP oracle."""
    import torch
    from mlir_hashjoin_b200 import datagen, join
    b = datagen.RelationSpec(1 << 20, 4, datagen.KIND_UNIQUE, 42, 0, 1 << 20)
    p = datagen.RelationSpec(1 << 24, 4, datagen.KIND_UNIFORM, 43, 0, 1 << 21)
    dR, dS = datagen.generate(b), datagen.generate(p)
    a, bb = join.hash_join(dR, dS)
    oa, ob = oracle.join(dR.cpu().numpy(), dS.cpu().numpy(), threads=0)
    assert a.numel() == oa.size
    assert join.pair_digest(a, bb) == oracle.pair_digest(oa, ob)
    # size-independent properties: every pair joins equal keys; probe rows are distinct (unique build)
    assert bool((dR[a.long()] == dS[bb.long()]).all())
    assert torch.unique(bb).numel() == bb.numel()


@pytest.mark.parametrize("kb,nR,nS", [(4, 6_000_000, 20_000_003), (8, 4_000_000, 12_000_003)])
@pytest.mark.parametrize("dups", [False, True])
def test_big_table_path(lib, cuda, oracle, layout, dups, kb, nR, nS):
    """Tables beyond L2 reach (> 48 MB of inline buckets): both relations are partitioned first (K5) — once, on the bucket hash, for one
    global table built and probed slice by slice; or twice, for the radix join in shared memory. Sparse build keys (odd multiplier, no dense range), i32 and i64, unique (inline layout) and 3x duplicated (grouped
    layout), with payload columns, against the OpenMP oracle: count + order-independent digest, plus key equality of every pair.
    Also: the same join with the big-table path switched off, and the reference's call sequence (countRows has no probe row ids;
    hash_join states them at count time through hjCountRows)."""
    import torch
    from mlir_hashjoin_b200 import datagen, join
    if layout in ("cache", "lists"):
        pytest.skip("sparse keys never take the direct-address layout: same code path as auto / hash-lists")
    dom = nR // 3 if dups else nR
    mul = 0x9E3779B1 if kb == 4 else datagen.ODD_MUL64
    b = datagen.RelationSpec(nR, kb, datagen.KIND_FK if dups else datagen.KIND_UNIQUE, 7, 0, dom, 0, mul)
    p = datagen.RelationSpec(nS, kb, datagen.KIND_UNIFORM, 8, 0, 2 * dom, 0, mul)
    dR, dS = datagen.generate(b), datagen.generate(p)
    pr = torch.arange(nR, dtype=torch.int32, device=cuda) * 3 + 1
    ps = torch.arange(nS, dtype=torch.int32, device=cuda) + 77
    a, bb = join.hash_join(dR, dS, buildPayload=pr, probePayload=ps)
    R, S = dR.cpu().numpy(), dS.cpu().numpy()
    oa, ob = oracle.join(R, S, threads=0)
    assert a.numel() == oa.size
    assert join.pair_digest(a, bb) == oracle.pair_digest(oa * 3 + 1, ob + 77)
    assert bool((dR[((a - 1) // 3).long()] == dS[(bb - 77).long()]).all())
    lib.hjSetSliced(0)                                             # same join with the radix layout (two partition passes, shared-memory tables)
    try:
        ar, br = join.hash_join(dR, dS, buildPayload=pr, probePayload=ps)
        tr = join.allocateHashTable(nR, None, dR.dtype, cuda)
        join.buildTable(dR, tr, pr)
        assert lib.hjTableLayout(tr.storage.data_ptr(), None) == 3
        del tr
    finally:
        lib.hjSetSliced(1)
    assert join.pair_digest(ar, br) == join.pair_digest(a, bb)
    lib.hjSetLocality(0)                                           # same join in input order: identical multiset
    a2, b2 = join.hash_join(dR, dS, buildPayload=pr, probePayload=ps)
    lib.hjSetLocality(1)
    assert join.pair_digest(a2, b2) == join.pair_digest(a, bb)
    # the reference's call sequence: row ids only at write time (payload column), and none at all (row base)
    table = join.allocateHashTable(nR, None, dR.dtype, cuda)
    join.buildTable(dR, table, pr)
    # one table built in table-slice order: inline for unique keys, grouped for duplicates (found by the sample, or — sample off — by
    # the inline attempt)
    assert lib.hjTableLayout(table.storage.data_ptr(), None) == (0x202 if dups else 0x200)
    n = join.countRows(dS, table)
    a3 = torch.empty(n, dtype=torch.int32, device=cuda); b3 = torch.empty(n, dtype=torch.int32, device=cuda)
    join.probeRelation(dS, table, a3, b3, ps)
    assert n == a.numel() and join.pair_digest(a3, b3) == join.pair_digest(a, bb)
    join.buildTable(dR, table, None, 5)
    n = join.countRows(dS, table)
    assert n == oa.size
    join.probeRelation(dS, table, a3, b3, None, 9)
    assert join.pair_digest(a3, b3) == oracle.pair_digest(oa + 5, ob + 9)


def test_c4_shape_fk_zipf_i64(lib, cuda, oracle, layout):
    """BASELINE.json config 4 at 1/8 scale: int64 keys, 2^22 build rows with every key exactly 4 times (FK 1:N), 2^24
    Zipf(1.0)-skewed probe keys -> 2^26 pairs (output-materialisation stress, hot keys). Count + digest vs the oracle."""
    import torch
    from mlir_hashjoin_b200 import datagen, join
    if layout in ("cache", "lists", "hash-lists"):
        pytest.skip("duplicate sparse keys: same code path as auto / hash")
    cfg = datagen.shrink(datagen.config("C4"), 1 << 22, 1 << 24)
    dR, dS = datagen.generate(cfg.build), datagen.generate(cfg.probe)
    a, bb = join.hash_join(dR, dS)
    assert a.numel() == 4 << 24
    oa, ob = oracle.join(dR.cpu().numpy(), dS.cpu().numpy(), threads=0)
    assert a.numel() == oa.size
    assert join.pair_digest(a, bb) == oracle.pair_digest(oa, ob)
    assert bool((dR[a.long()] == dS[bb.long()]).all())


@pytest.mark.slow
def test_c2_full_size_properties(lib, cuda):
    """BASELINE.json config 2 at full size (2^24 x 2^28): analytic count, key equality of every pair, probe rows a permutation."""
    import torch
    from mlir_hashjoin_b200 import datagen, join
    cfg = datagen.config("C2")
    dR, dS = datagen.generate(cfg.build), datagen.generate(cfg.probe)
    a, b = join.hash_join(dR, dS)
    assert a.numel() == cfg.expected_out
    assert bool((dR[a.long()] == dS[b.long()]).all())
    bs, _ = torch.sort(b)
    assert bool((bs == torch.arange(cfg.probe.n, dtype=torch.int32, device=cuda)).all())


def test_cpp_host_driver_main(lib, cuda):
    """The C++ emulation of the reference's lowered @main (join_v1.mlir:525-649) calling the library by symbol name:
    four timer lines, the result size, and check()'s verdict 1 — at the .ll snapshot shape and at a larger one."""
    import os
    import re
    import subprocess
    from mlir_hashjoin_b200.build import DRIVER
    env = dict(os.environ, HASHJOIN_SEED_R="11", HASHJOIN_SEED_S="12")
    for argv in ([], ["12", "12", "5", "4"], ["100000", "400000", "1000", "50000"]):
        r = subprocess.run([str(DRIVER), *argv], capture_output=True, text=True, env=env, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        vals = re.findall(r"^\[(-?\d+)\]$", r.stdout, flags=re.M)
        assert len(vals) == 2 and vals[1] == "1", r.stdout
        # init, build, count always print; probe only when the result is not empty (join_v1.mlir:600-601)
        assert len(re.findall(r"For \d+, time taken: \d+ microseconds", r.stdout)) == (4 if int(vals[0]) else 3)
    assert int(vals[0]) > 0


def test_partition_push_single_gpu(lib, cuda, oracle):
    """K5 fused with the exchange, on one GPU: hjPartitionCount + hjPartitionPush with every "peer" buffer in local memory (the kernel
    cannot tell). Every tuple must land in its owner's buffer, inside the region the cursors gave this "rank", keys and row ids together."""
    import torch
    rng = np.random.default_rng(18)
    for kd, kb, n, parts in ((np.int64, 8, 300_007, 2), (np.int32, 4, 1_000_003, 8), (np.int64, 8, 70_001, 5), (np.int64, 8, 3_000_017, 8)):
        keys = rng.integers(-2**40, 2**40, n).astype(kd)
        d = torch.from_numpy(keys).to(cuda)
        counts = torch.empty(parts, dtype=torch.int64, device=cuda)
        ws = torch.empty(lib.hjPartitionWorkspaceBytes(n, parts), dtype=torch.uint8, device=cuda)
        assert lib.hjPartitionCount(d.data_ptr(), n, kb, parts, counts.data_ptr(), ws.data_ptr(), ws.numel(), None) == 0
        cnt = counts.cpu().numpy()
        assert cnt.sum() == n
        lead = 1000                                                             # pretend lower ranks already own the first 1000 slots of every buffer
        bufs_k = [torch.full((lead + int(c) + 64,), -7, dtype=d.dtype, device=cuda) for c in cnt]
        bufs_r = [torch.full((lead + int(c) + 64,), -7, dtype=torch.int32, device=cuda) for c in cnt]
        kp = torch.tensor([b.data_ptr() for b in bufs_k], dtype=torch.int64, device=cuda)
        rp = torch.tensor([b.data_ptr() for b in bufs_r], dtype=torch.int64, device=cuda)
        cursors = torch.full((parts,), lead, dtype=torch.int64, device=cuda)
        assert lib.hjPartitionPush(d.data_ptr(), None, 5, n, kb, parts, kp.data_ptr(), rp.data_ptr(), cursors.data_ptr(), ws.data_ptr(), ws.numel(), None) == 0
        torch.cuda.synchronize()
        seen = np.zeros(n, dtype=bool)
        for p in range(parts):
            bk, br = bufs_k[p].cpu().numpy(), bufs_r[p].cpu().numpy()
            assert (bk[:lead] == -7).all() and (bk[lead + cnt[p]:] == -7).all() and (br[:lead] == -7).all() and (br[lead + cnt[p]:] == -7).all()
            rows = br[lead:lead + cnt[p]] - 5
            assert np.array_equal(keys[rows], bk[lead:lead + cnt[p]])          # rows travelled with their keys
            assert not seen[rows].any()
            seen[rows] = True
        assert seen.all()
