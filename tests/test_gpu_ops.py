"""GPU parity of the operators either side of the join path against the oracle: late gather and row
materialisation (nested-loop.mlir:165-187), selection (Experiments/selection.mlir:34-155), semi-join / count-only, two-column
keys, and the bounded device window of hjJoinHost (out-of-core probe side). All through the C ABI; bit-exact."""
import numpy as np
import pytest

from oracle.binding import sorted_pairs

pytestmark = pytest.mark.gpu


def _sorted_rows(a):
    a = np.asarray(a)
    return a[np.lexsort(a.T[::-1])] if a.shape[0] else a


def test_gather_columns(lib, cuda, oracle):
    import torch
    from mlir_hashjoin_b200 import join
    rng = np.random.default_rng(3)
    for dt, n_col, n in ((np.int32, 100_003, 1_000_001), (np.int64, 7, 5000), (np.float32, 50_000, 3), (np.float64, 1 << 20, (1 << 21) + 5)):
        col = (rng.integers(-2**31, 2**31, n_col)).astype(dt)
        rows = rng.integers(0, n_col, n).astype(np.int32)
        got = join.gather(torch.from_numpy(col).to(cuda), torch.from_numpy(rows).to(cuda))
        assert np.array_equal(got.cpu().numpy(), col[rows])
        got = join.gather(torch.from_numpy(col).to(cuda), torch.from_numpy(rows + 1000).to(cuda), rowBase=1000)
        assert np.array_equal(got.cpu().numpy(), col[rows])
    empty = join.gather(torch.from_numpy(col).to(cuda), torch.empty(0, dtype=torch.int32, device=cuda))
    assert empty.numel() == 0
    # unaligned row-id pointer (memref offset): the scalar path gives the same result
    r = torch.from_numpy(rows).to(cuda)[1:]
    assert np.array_equal(join.gather(torch.from_numpy(col).to(cuda), r).cpu().numpy(), col[rows[1:]])


def test_join_then_gather_payload_columns(lib, cuda, oracle):
    """The step right after the path: joined payload columns from the pair stream (build payload i64, probe payload f32)."""
    import torch
    from mlir_hashjoin_b200 import join
    rng = np.random.default_rng(5)
    R = rng.integers(0, 5000, 20000).astype(np.int32); S = rng.integers(0, 5000, 60001).astype(np.int32)
    payR = rng.integers(-2**62, 2**62, R.size).astype(np.int64); payS = rng.random(S.size).astype(np.float32)
    a, b = join.hash_join(torch.from_numpy(R).to(cuda), torch.from_numpy(S).to(cuda))
    gR = join.gather(torch.from_numpy(payR).to(cuda), a).cpu().numpy(); gS = join.gather(torch.from_numpy(payS).to(cuda), b).cpu().numpy()
    oa, ob = oracle.join(R, S)
    got = np.stack([a.cpu().numpy().astype(np.int64), b.cpu().numpy().astype(np.int64), gR, gS.view(np.int32).astype(np.int64)], axis=1)
    want = np.stack([oa.astype(np.int64), ob.astype(np.int64), payR[oa], payS[ob].view(np.int32).astype(np.int64)], axis=1)
    assert np.array_equal(_sorted_rows(got), _sorted_rows(want))


def test_nested_loop_kat_and_row_materialisation(lib, cuda, oracle):
    """nested-loop.mlir:208-212 fixed tables (val = i + j, 20 x 3 and 20 x 2): 20 result rows [i, i+1, i+2, i+1]; then random tables with
    duplicate keys on both sides, either side larger, against the oracle's restatement of :78-188 (rows compared as a sorted multiset)."""
    import torch
    from mlir_hashjoin_b200 import join
    t1 = (np.arange(20)[:, None] + np.arange(3)[None, :]).astype(np.int32)
    t2 = (np.arange(20)[:, None] + np.arange(2)[None, :]).astype(np.int32)
    got = join.nested_loop_join(torch.from_numpy(t1).to(cuda), torch.from_numpy(t2).to(cuda)).cpu().numpy()
    want = np.stack([np.arange(20), np.arange(20) + 1, np.arange(20) + 2, np.arange(20) + 1], axis=1).astype(np.int32)
    assert np.array_equal(_sorted_rows(got), want)
    assert np.array_equal(_sorted_rows(oracle.nested_join_rows(t1, t2)), want)
    rng = np.random.default_rng(8)
    for (r1, c1, r2, c2, dom) in ((300, 4, 1000, 3, 200), (1500, 2, 400, 5, 5000), (1, 1, 1, 1, 1), (64, 3, 64, 1, 10)):
        a = rng.integers(0, dom, (r1, c1)).astype(np.int32); b = rng.integers(0, dom, (r2, c2)).astype(np.int32)
        got = join.nested_loop_join(torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)).cpu().numpy()
        x, y = (b, a) if r1 < r2 else (a, b)                       # the larger table is the outer side (nested-loop.mlir:252-262)
        want = oracle.nested_join_rows(x, y)
        assert got.shape == want.shape and np.array_equal(_sorted_rows(got), _sorted_rows(want))


def test_selection_matches_oracle(lib, cuda, oracle):
    """Experiments/selection.mlir: `value < 80.0` over f32 (:52,62), plus the other comparisons / types; NaN never passes an ordered
    comparison; output in input order, sizes around the 512-row warp slices."""
    import torch
    from mlir_hashjoin_b200 import join
    rng = np.random.default_rng(12)
    col = (rng.random(100_003) * 160).astype(np.float32)
    col[::97] = np.nan
    v, r = join.selection(torch.from_numpy(col).to(cuda), "<", 80.0)
    ov, orow = oracle.select(col, 0, 80.0)
    assert np.array_equal(v.cpu().numpy(), ov) and np.array_equal(r.cpu().numpy(), orow) and 40_000 < ov.size < 60_000
    for dt, c in ((np.int32, 17), (np.int64, -3), (np.float64, 0.25), (np.float32, 5.0)):
        for n in (0, 1, 511, 512, 513, 70_001):
            x = (rng.integers(-50, 50, n) if np.issubdtype(dt, np.integer) else rng.normal(size=n) * 4).astype(dt)
            for op, code in join.SELECT_OPS.items():
                v, r = join.selection(torch.from_numpy(x).to(cuda), op, c, rowBase=7)
                ov, orow = oracle.select(x, code, c)
                assert np.array_equal(v.cpu().numpy(), ov) and np.array_equal(r.cpu().numpy(), orow + 7), (dt, n, op)


@pytest.mark.parametrize("policy", ["auto", "hash", "lists"])
def test_semi_join_and_count_only(lib, cuda, oracle, policy):
    """Key-only mode: probe rows with >= 1 match, each once, on every layout (direct-address, bucketised hash, grouped, radix) and both
    unique-layout probe paths; count-only = hjCount without a write pass."""
    import torch
    from mlir_hashjoin_b200 import datagen, join
    lib.hjSetAllowDense({"auto": 2, "hash": 0, "lists": 1}[policy]); lib.hjSetSparse(2 if policy == "lists" else 1)
    try:
        rng = np.random.default_rng(21)
        cases = [(rng.permutation(50_000)[:30_000].astype(np.int32), rng.integers(0, 80_000, 200_003).astype(np.int32)),        # unique build
                 (rng.integers(0, 3000, 40_000).astype(np.int32), rng.integers(0, 6000, 100_001).astype(np.int32)),            # duplicate build keys
                 (rng.integers(0, 3000, 40_000).astype(np.int64) * 0x9E3779B97F4A7, rng.integers(0, 6000, 100_001).astype(np.int64) * 0x9E3779B97F4A7)]
        for R, S in cases:
            dR, dS = torch.from_numpy(R).to(cuda), torch.from_numpy(S).to(cuda)
            table = join.allocateHashTable(R.size, None, dR.dtype, cuda)
            join.buildTable(dR, table)
            got = join.semi_join(dS, table, probeRowBase=5).cpu().numpy()
            oa, ob = oracle.join(R, S)
            assert np.array_equal(np.sort(got), np.unique(ob) + 5)
            assert join.countRows(dS, table) == oa.size                      # count-only: the size without writing anything
        # radix layout (table beyond L2 reach), 3x duplicated sparse i64 keys
        b = datagen.RelationSpec(1_800_000, 8, datagen.KIND_FK, 7, 0, 600_000, 0, datagen.ODD_MUL64)
        p = datagen.RelationSpec(3_000_001, 8, datagen.KIND_UNIFORM, 8, 0, 1_200_000, 0, datagen.ODD_MUL64)
        dR, dS = datagen.generate(b), datagen.generate(p)
        oa, ob = oracle.join(dR.cpu().numpy(), dS.cpu().numpy(), threads=0)
        for sliced, want_layout in ((0, 3), (1, 0x202)):                  # radix layout, then the slice-ordered grouped table
            lib.hjSetSliced(sliced)
            table = join.allocateHashTable(b.n, None, dR.dtype, cuda)
            join.buildTable(dR, table)
            assert lib.hjTableLayout(table.storage.data_ptr(), None) == want_layout
            got = join.semi_join(dS, table).cpu().numpy()
            assert np.array_equal(np.sort(got), np.unique(ob))
    finally:
        lib.hjSetAllowDense(2); lib.hjSetSparse(1); lib.hjSetSliced(1)


def test_two_column_keys(lib, cuda):
    """Multi-column equi-join (projectDescription.md:28): pack (a, b) into one i64 key; pairs must be exactly those with a == a' and b == b'."""
    import torch
    from mlir_hashjoin_b200 import join
    rng = np.random.default_rng(31)
    a1 = rng.integers(-5, 5, 2000).astype(np.int32); b1 = rng.integers(-40, 40, 2000).astype(np.int32)
    a2 = rng.integers(-5, 5, 3000).astype(np.int32); b2 = rng.integers(-40, 40, 3000).astype(np.int32)
    k1 = join.pack_keys(torch.from_numpy(a1).to(cuda), torch.from_numpy(b1).to(cuda))
    k2 = join.pack_keys(torch.from_numpy(a2).to(cuda), torch.from_numpy(b2).to(cuda))
    outR, outS = join.hash_join(k1, k2)
    i, j = np.nonzero((a1[:, None] == a2[None, :]) & (b1[:, None] == b2[None, :]))
    assert np.array_equal(sorted_pairs(outR.cpu().numpy(), outS.cpu().numpy()), sorted_pairs(i, j))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_float_keys(lib, cuda, dt):
    """Non-integer keys (projectDescription.md:28): float columns encoded by hjEncodeFloatKeys join exactly where IEEE `==` holds —
    -0.0 with +0.0, NaN with nothing (not even the same NaN), infinities and denormals with themselves."""
    import torch
    from mlir_hashjoin_b200 import join
    rng = np.random.default_rng(53)
    special = np.array([0.0, -0.0, np.nan, -np.nan, np.inf, -np.inf, np.finfo(dt).tiny / 4, 1.5, -1.5], dtype=dt)
    R = np.concatenate([special, rng.integers(-50, 50, 3000).astype(dt) / 4]).astype(dt)
    S = np.concatenate([rng.integers(-60, 60, 5000).astype(dt) / 4, special, special]).astype(dt)
    rng.shuffle(R); rng.shuffle(S)
    kR = join.encode_float_keys(torch.from_numpy(R).to(cuda), probe_side=False)
    kS = join.encode_float_keys(torch.from_numpy(S).to(cuda), probe_side=True)
    assert kR.dtype == (torch.int32 if dt == np.float32 else torch.int64)
    outR, outS = join.hash_join(kR, kS)
    i, j = np.nonzero(R[:, None] == S[None, :])                       # numpy's == is IEEE: NaN != NaN, -0.0 == 0.0
    assert np.array_equal(sorted_pairs(outR.cpu().numpy(), outS.cpu().numpy()), sorted_pairs(i, j))


def test_join_host_bounded_window(lib, cuda, oracle):
    """hjJoinHost keeps only three probe chunks and two result slots on the device: many small chunks (ring wrap-around, growing result
    slots, duplicate build keys so chunk results differ in size) give the oracle's multiset."""
    import torch
    from mlir_hashjoin_b200 import join
    rng = np.random.default_rng(41)
    R = rng.integers(0, 20_000, 30_000).astype(np.int32); S = rng.integers(0, 40_000, 1_000_003).astype(np.int32)
    lib.hjSetHostChunkRows(1 << 16)
    try:
        a, b, n = join.join_host(torch.from_numpy(R), torch.from_numpy(S))
    finally:
        lib.hjSetHostChunkRows(1 << 24)
    oa, ob = oracle.join(R, S, threads=0)
    assert n == oa.size and oracle.pair_digest(a.numpy(), b.numpy()) == oracle.pair_digest(oa, ob)
    assert np.array_equal(R[a.numpy()], S[b.numpy()])                          # every pair joins equal keys


def test_per_table_policy_and_concurrent_threads(lib, cuda, oracle):
    """The policy is a property of the TABLE (hjBuildEx freezes it into the device-resident header), not of the process: two tables built
    under different policies coexist and each is probed by its own rules whatever hjSet* says later. And the library keeps no per-call
    state on the host: two host threads drive their own (table, scratch, stream) concurrently and both get the oracle's result."""
    import threading
    import torch
    from mlir_hashjoin_b200 import _lib
    rng = np.random.default_rng(51)
    R = rng.permutation(200_000).astype(np.int32)[:120_000]                 # unique keys in a dense range
    S = rng.integers(0, 250_000, 900_001).astype(np.int32)
    dR, dS = torch.from_numpy(R).to(cuda), torch.from_numpy(S).to(cuda)
    oa, ob = oracle.join(R, S)
    want = sorted_pairs(oa, ob)
    tb, sb = lib.hjTableBytes(R.size, 4), lib.hjScratchBytes(S.size, 4)
    POL_HASH = 0 | (1 << 2) | (1 << 3) | (1 << 5)                            # HJ_POLICY_DENSE_OFF | RADIX | LISTS_SAMPLED | DUP_SAMPLE
    POL_DENSE = 2 | (1 << 2) | (2 << 3) | (1 << 5)                           # HJ_POLICY_DENSE_RANGE | RADIX | LISTS_ALWAYS | DUP_SAMPLE
    tables = [torch.empty(tb, dtype=torch.uint8, device=cuda) for _ in range(2)]
    assert lib.hjBuildEx(dR.data_ptr(), R.size, 4, None, 0, tables[0].data_ptr(), tb, POL_HASH, None) == 0
    assert lib.hjBuildEx(dR.data_ptr(), R.size, 4, None, 0, tables[1].data_ptr(), tb, POL_DENSE, None) == 0
    lib.hjSetAllowDense(1); lib.hjSetSparse(0)                               # process defaults changed AFTER the builds: must not matter
    try:
        assert lib.hjTableLayout(tables[0].data_ptr(), None) == 0 and lib.hjTableLayout(tables[1].data_ptr(), None) & 0xFF == 1
        results, errors = [None, None], []

        def worker(k):
            try:
                with torch.cuda.stream(torch.cuda.Stream(device=cuda)) as _:
                    st = torch.cuda.current_stream().cuda_stream
                    scratch = torch.empty(sb, dtype=torch.uint8, device=cuda)
                    for _rep in range(20):                                   # many small calls: the readbacks of the two threads interleave
                        n = lib.hjCount(dS.data_ptr(), S.size, 4, tables[k].data_ptr(), scratch.data_ptr(), sb, st)
                        outR = torch.empty(n, dtype=torch.int32, device=cuda); outS = torch.empty(n, dtype=torch.int32, device=cuda)
                        _lib.check_status(lib.hjWrite(dS.data_ptr(), S.size, 4, tables[k].data_ptr(), scratch.data_ptr(), outR.data_ptr(), outS.data_ptr(), None, 0, st), "hjWrite")
                        torch.cuda.current_stream().synchronize()
                        assert n == oa.size
                    results[k] = (outR.cpu().numpy(), outS.cpu().numpy(), lib.hjProbePath(scratch.data_ptr(), S.size, 4, st))
            except Exception as e:                                           # noqa: BLE001
                errors.append(e)
        threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        for k in range(2):
            assert np.array_equal(sorted_pairs(results[k][0], results[k][1]), want)
        assert results[0][2] == 0 and results[1][2] == 1                     # table 0: sampled (48 % hit -> match cache); table 1: hit lists always
    finally:
        lib.hjSetAllowDense(2); lib.hjSetSparse(1)
