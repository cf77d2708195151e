"""torchrun --nproc-per-node N tools/dist_check.py — multi-GPU parity: both plans of mlir-hashjoin_b200/dist.py against
the oracle on seeded inputs (digest + count of the global result; sorted compare at this size too)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from mlir_hashjoin_b200 import datagen, dist as hjdist
    from oracle import Oracle
    from oracle.binding import sorted_pairs
    o = Oracle()
    ok = True
    for name, b, p in (
        ("i32 unique x uniform", datagen.RelationSpec(200_003, 4, datagen.KIND_UNIQUE, 42, 0, 300_000), datagen.RelationSpec(1_000_003, 4, datagen.KIND_UNIFORM, 43, 0, 400_000)),
        ("i64 fk x zipf", datagen.RelationSpec(1 << 16, 8, datagen.KIND_FK, 46, 0, 1 << 14, 0, datagen.ODD_MUL64), datagen.RelationSpec(300_001, 8, datagen.KIND_ZIPF, 47, 0, 1 << 14, 0, datagen.ODD_MUL64)),
        # local joins beyond L2 reach on every rank: the radix layout behind both multi-GPU plans
        ("i64 unique big", datagen.RelationSpec(4_000_003 * world, 8, datagen.KIND_UNIQUE, 48, 0, 4_000_003 * world, 0, datagen.ODD_MUL64),
         datagen.RelationSpec(6_000_001 * world, 8, datagen.KIND_UNIFORM, 49, 0, 6_000_000 * world, 0, datagen.ODD_MUL64)),
    ):
        R = o.generate(b.n, b.key_bytes, b.kind, b.seed, b.lo, b.domain, b.p16, b.key_mul)
        S = o.generate(p.n, p.key_bytes, p.kind, p.seed, p.lo, p.domain, p.p16, p.key_mul)
        want = sorted_pairs(*o.join(R, S, threads=0))
        blo, bhi = hjdist.shard_range(b.n, rank, world); plo, phi = hjdist.shard_range(p.n, rank, world)
        dS = datagen.generate(p, dev, plo, phi - plo)
        # plan 1: broadcast build (build relation born on rank 0)
        dR = datagen.generate(b, dev) if rank == 0 else torch.empty(b.n, dtype=b.dtype, device=dev)
        a1, b1 = hjdist.broadcast_build_join(dR, dS, plo)
        # plan 2: radix partition + all-to-all (both relations range-sharded)
        dRs = datagen.generate(b, dev, blo, bhi - blo)
        a2, b2 = hjdist.radix_join(dRs, blo, dS, plo)
        # plan 3: same, exchange fused into the partition kernel (peer stores over NVLink into symmetric buffers)
        bx = hjdist.PeerExchange(int(1.5 * b.n / world) + 4096, b.dtype, dev)
        px = hjdist.PeerExchange(int(1.5 * p.n / world) + 4096, p.dtype, dev)
        a3, b3 = hjdist.radix_join_fused(dRs, blo, dS, plo, bx, px)
        a3, b3 = a3.clone(), b3.clone()
        a4, b4 = hjdist.radix_join_fused(dRs, blo, dS, plo, bx, px)      # buffers are reusable step after step
        a4, b4 = a4.clone(), b4.clone()
        # plan 4: parts staged locally, carried by the copy engines beside the partition / build kernels (twice: staging reused)
        a6, b6 = hjdist.radix_join_staged(dRs, blo, dS, plo, bx, px)
        a6, b6 = a6.clone(), b6.clone()
        a7, b7 = hjdist.radix_join_staged(dRs, blo, dS, plo, bx, px)
        # receive buffers too small for the key distribution: every rank raises alike, nothing was stored, and the NCCL all-to-all
        # plan (exact sizes) gives the oracle's result
        from mlir_hashjoin_b200._lib import HashJoinError
        tiny_b = hjdist.PeerExchange(max(16, b.n // (4 * world)), b.dtype, dev); tiny_p = hjdist.PeerExchange(max(16, p.n // (4 * world)), p.dtype, dev)
        try:
            hjdist.radix_join_fused(dRs, blo, dS, plo, tiny_b, tiny_p)
            overflow_raised = False
        except HashJoinError:
            overflow_raised = True
        a5, b5 = hjdist.radix_join(dRs, blo, dS, plo)
        if rank == 0:
            print(f"{name:22s} overflow   world={world} raised={overflow_raised} parity={'OK' if overflow_raised else 'FAIL'}", flush=True)
            ok = ok and overflow_raised
        for plan, (a, bb) in (("broadcast", (a1, b1)), ("radix", (a2, b2)), ("radix-fused", (a3, b3)), ("radix-fused#2", (a4, b4)), ("radix-staged", (a6, b6)), ("radix-staged#2", (a7, b7)),
                               ("radix-after-overflow", (a5, b5))):
            n = torch.tensor([a.numel()], device=dev)
            sizes = [torch.zeros_like(n) for _ in range(world)]
            dist.all_gather(sizes, n)
            cap = int(max(s.item() for s in sizes))
            pad = torch.full((2, cap), -1, dtype=torch.int32, device=dev)
            pad[0, :a.numel()] = a; pad[1, :a.numel()] = bb
            allp = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(allp, pad)
            if rank == 0:
                got = np.concatenate([allp[r][:, :int(sizes[r].item())].cpu().numpy() for r in range(world)], axis=1)
                same = got.shape[1] == want.shape[0] and np.array_equal(sorted_pairs(got[0], got[1]), want)
                print(f"{name:22s} {plan:10s} world={world} pairs={got.shape[1]} parity={'OK' if same else 'FAIL'}", flush=True)
                ok = ok and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
