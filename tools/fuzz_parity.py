"""Randomised parity run: the CUDA path through the C ABI against the oracle on random shapes — sizes around the layout thresholds
(direct-address / inline / grouped / slice-ordered / radix), both key widths, key distributions from unique to heavily duplicated, dense
and sparse key ranges, payload columns and row bases, every layout policy. Count + order-independent pair digest (+ key equality of
every pair); stops at the first mismatch with the case's parameters. Checker infrastructure: uses oracle/ like the tests do.

    python tools/fuzz_parity.py [--seconds 180] [--seed 1] [--max-rows 30000000]"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main() -> None:
    import torch
    from mlir_hashjoin_b200 import _lib, datagen, join
    from oracle import Oracle

    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=180.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-rows", type=int, default=30_000_000)
    args = ap.parse_args()
    lib, o = _lib.load(), Oracle()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(args.seed)
    t0, cases, layouts = time.time(), 0, {}
    sizes = [1, 2, 31, 257, 4097, 70_001, 1_000_003, 3_000_017, 6_500_000, 9_000_001, 17_000_003, args.max_rows]
    while time.time() - t0 < args.seconds:
        kb = int(rng.choice([4, 8]))
        nR = int(min(args.max_rows, rng.choice(sizes) * rng.uniform(0.5, 1.5))) or 1
        nS = int(min(args.max_rows, rng.choice(sizes) * rng.uniform(0.5, 1.5)))
        kind = int(rng.choice([datagen.KIND_UNIQUE, datagen.KIND_FK, datagen.KIND_UNIFORM]))
        dup = int(rng.choice([1, 1, 2, 4, 50, 2000]))
        dom = max(1, nR // dup) if kind != datagen.KIND_UNIQUE else nR
        mul = 0 if rng.random() < 0.4 else (0x9E3779B1 if kb == 4 else datagen.ODD_MUL64)
        lo = int(rng.choice([0, 0, -1000, 12345]))
        b = datagen.RelationSpec(nR, kb, kind, int(rng.integers(1, 1 << 30)), lo, dom, 0, mul)
        p = datagen.RelationSpec(nS, kb, int(rng.choice([datagen.KIND_UNIFORM, datagen.KIND_FK])), int(rng.integers(1, 1 << 30)), lo, max(1, int(dom * rng.choice([1, 1, 2, 10]))), 0, mul)
        policy = {"dense": int(rng.choice([0, 1, 2, 2])), "sparse": int(rng.choice([0, 1, 1, 2])), "sliced": int(rng.choice([0, 1, 1])), "locality": int(rng.choice([0, 1, 1, 1])),
                  "dupsample": int(rng.choice([0, 1, 1])), "threads": int(rng.choice([0, 0, 256, 512, 1024]))}
        payload = bool(rng.random() < 0.5)
        case = {"kb": kb, "nR": nR, "nS": nS, "build": [b.kind, b.seed, b.lo, b.domain, b.key_mul], "probe": [p.kind, p.seed, p.lo, p.domain, p.key_mul], "policy": policy, "payload": payload}
        lib.hjSetAllowDense(policy["dense"]); lib.hjSetSparse(policy["sparse"]); lib.hjSetSliced(policy["sliced"]); lib.hjSetLocality(policy["locality"])
        lib.hjSetDupSample(policy["dupsample"]); lib.hjSetPartitionThreads(policy["threads"])
        try:
            dR, dS = datagen.generate(b, dev), datagen.generate(p, dev)
            R, S = dR.cpu().numpy(), dS.cpu().numpy()
            oa, ob = o.join(R, S, threads=0)
            if oa.size > 600_000_000:
                continue
            if payload:
                pr = torch.arange(nR, dtype=torch.int32, device=dev) * 2 + 3
                ps = torch.arange(nS, dtype=torch.int32, device=dev) + 11
                a, bb = join.hash_join(dR, dS, buildPayload=pr, probePayload=ps)
                want = o.pair_digest(oa * 2 + 3, ob + 11)
            else:
                a, bb = join.hash_join(dR, dS)
                want = o.pair_digest(oa, ob)
            table = join.allocateHashTable(nR, None, dR.dtype, dev)
            join.buildTable(dR, table)
            code = lib.hjTableLayout(table.storage.data_ptr(), None)
            layouts[hex(code)] = layouts.get(hex(code), 0) + 1
            ok = a.numel() == oa.size and join.pair_digest(a, bb) == want
            if ok and a.numel() and not payload:
                ok = bool((dR[a.long()] == dS[bb.long()]).all())
            # semi-join and the reference's call sequence on the same table
            if ok:
                n = join.countRows(dS, table)
                a3 = torch.empty(n, dtype=torch.int32, device=dev); b3 = torch.empty(n, dtype=torch.int32, device=dev)
                if n:
                    join.probeRelation(dS, table, a3, b3)
                ok = n == oa.size and join.pair_digest(a3, b3) == o.pair_digest(oa, ob)
            if ok:
                rows = join.semi_join(dS, table)
                ok = rows.numel() == np.unique(ob).size and bool((torch.sort(rows).values.cpu() == torch.from_numpy(np.unique(ob).astype(np.int32))).all())
            del table
        finally:
            lib.hjSetAllowDense(2); lib.hjSetSparse(1); lib.hjSetSliced(1); lib.hjSetLocality(1); lib.hjSetDupSample(1); lib.hjSetPartitionThreads(0)
        cases += 1
        if not ok:
            print(json.dumps({"result": "MISMATCH", "case": case, "got": int(a.numel()), "want": int(oa.size)}))
            sys.exit(1)
    print(json.dumps({"result": "ok", "cases": cases, "seconds": round(time.time() - t0, 1), "seed": args.seed, "table_layouts_seen": layouts}))


if __name__ == "__main__":
    main()
