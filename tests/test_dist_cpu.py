"""CPU-only, world_size 2 over gloo: the host-side logic of the multi-GPU plans (row-range sharding, the count-matrix
all-gather, the variable-count all-to-all). The partition step itself is a CUDA kernel (K5, covered by the gpu tests);
here the test partitions with numpy so that the exchange plumbing in mlir-hashjoin_b200/dist.py runs on CPU tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _np_partition(keys, rows, world):
    part = (keys.astype(np.int64) * 2654435761 % (1 << 32) * world) >> 32        # any deterministic key -> rank map
    order = np.argsort(part, kind="stable")
    counts = np.bincount(part, minlength=world)
    offsets = np.concatenate([[0], np.cumsum(counts)])
    return keys[order], rows[order], offsets


def _worker(rank, world, port, nR, nS, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mlir_hashjoin_b200 import dist as hjdist
    from oracle import Oracle
    rng = np.random.default_rng(123)
    R = rng.integers(0, 3000, nR).astype(np.int32); S = rng.integers(0, 3000, nS).astype(np.int32)
    blo, bhi = hjdist.shard_range(nR, rank, world); plo, phi = hjdist.shard_range(nS, rank, world)
    got = {}
    for name, arr, lo, hi in (("b", R, blo, bhi), ("p", S, plo, phi)):
        k, r, off = _np_partition(arr[lo:hi], np.arange(lo, hi, dtype=np.int32), world)
        plan = hjdist.exchange_plan(torch.from_numpy(off), None)
        assert sum(plan.send_counts) == hi - lo
        counts = torch.from_numpy(np.diff(off).astype(np.int64))
        rows_ = [torch.empty_like(counts) for _ in range(world)]
        dist.all_gather(rows_, counts)
        land = hjdist.landing_plan(torch.stack(rows_), rank)
        assert land.received == sum(plan.recv_counts) and land.offsets == off.tolist()
        mk = hjdist.exchange(torch.from_numpy(k), plan).numpy(); mr = hjdist.exchange(torch.from_numpy(r), plan).numpy()
        assert mk.size == sum(plan.recv_counts) == mr.size
        assert np.array_equal(arr[mr], mk)                                       # rows still carry their keys
        pk, _, _ = _np_partition(mk, mr, world)
        assert np.all(((mk.astype(np.int64) * 2654435761 % (1 << 32) * world) >> 32) == rank)   # only my partition arrived
        got[name] = (mk, mr)
    o = Oracle()
    a, b = o.join(got["b"][0], got["p"][0])
    np.save(os.path.join(out_dir, f"pairs_{rank}.npy"), np.stack([got["b"][1][a], got["p"][1][b]], axis=1))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range():
    from mlir_hashjoin_b200 import dist as hjdist
    for n in (0, 1, 7, 8, 1000, 2_000_000_000):
        for w in (1, 2, 3, 8):
            r = [hjdist.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_landing_plan_tiles_every_owner_buffer():
    """Peer-store / copy-engine exchanges: the regions the ranks claim in an owner's receive buffer tile [0, received) exactly, in rank
    order, and every rank computes the same capacity verdict from the same count matrix."""
    from mlir_hashjoin_b200 import dist as hjdist
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 8):
        for _ in range(5):
            m = torch.from_numpy(rng.integers(0, 1000, (world, world)) * rng.integers(0, 2, (world, world)))      # some empty parts
            plans = [hjdist.landing_plan(m, r) for r in range(world)]
            assert len({p.fullest for p in plans}) == 1 and plans[0].fullest == int(m.sum(0).max())
            for owner in range(world):
                regions = sorted((plans[src].first_at_owner[owner], int(m[src, owner])) for src in range(world))
                end = 0
                for first, count in regions:
                    assert first == end or count == 0
                    end = max(end, first + count)
                assert end == plans[owner].received == int(m[:, owner].sum())
            for src in range(world):
                assert plans[src].offsets[0] == 0 and plans[src].offsets[-1] == int(m[src].sum())
                assert [b - a for a, b in zip(plans[src].offsets, plans[src].offsets[1:])] == m[src].tolist()


@pytest.mark.timeout(180)
def test_radix_exchange_two_ranks_gloo(tmp_path, oracle):
    from oracle.binding import sorted_pairs
    world, nR, nS = 2, 5001, 12007
    mp.spawn(_worker, args=(world, _free_port(), nR, nS, str(tmp_path)), nprocs=world, join=True)
    pairs = np.concatenate([np.load(tmp_path / f"pairs_{r}.npy") for r in range(world)])
    rng = np.random.default_rng(123)
    R = rng.integers(0, 3000, nR).astype(np.int32); S = rng.integers(0, 3000, nS).astype(np.int32)
    a, b = oracle.join(R, S)
    assert np.array_equal(sorted_pairs(pairs[:, 0], pairs[:, 1]), sorted_pairs(a, b))
