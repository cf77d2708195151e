"""What the box gives every rank at once between pinned host memory and its GPU (the ceiling of bench.py's e2e leg at N GPUs).

Every rank binds like bench.py (CPUs and memory policy near its GPU), pins 1 GiB in and 1 GiB out, and after a barrier copies
host -> device, device -> host and both at the same time (two streams), all ranks together. Rank 0 prints one JSON line with the GB/s of
every rank. Run: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py [--no-bind]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main() -> None:
    import torch
    import torch.distributed as dist
    import bench

    ap = argparse.ArgumentParser()
    ap.add_argument("--no-bind", action="store_true")
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    note = "unbound (asked)" if args.no_bind else bench.bind_near_gpu(torch, dev)
    n = args.mib << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out.fill_(2)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(h2d: bool, d2h: bool) -> float:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a); s2.wait_event(a)
        for _ in range(args.reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        b.record(); b.synchronize()
        return a.elapsed_time(b) / args.reps

    timed(True, True)
    res = {"h2d_alone": n / timed(True, False) / 1e6, "d2h_alone": n / timed(False, True) / 1e6}
    both = timed(True, True)
    res["both_each_direction"] = n / both / 1e6
    res["bind"] = note
    out = [None] * world
    if world > 1:
        dist.all_gather_object(out, res)
    else:
        out = [res]
    if rank == 0:
        agg = {k: round(sum(r[k] for r in out), 1) for k in ("h2d_alone", "d2h_alone", "both_each_direction")}
        print(json.dumps({"ranks": world, "mib": args.mib, "aggregate_gbs": agg,
                          "per_rank_gbs": [{k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()} for r in out]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
