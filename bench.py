#!/usr/bin/env python
"""bench.py — join tuples/s (build + probe) for the B200 hash join, one JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c2s|c3|c4|c5|ref10m|ref100m] [--impl reference]

A step is one pass of the hot path over one batch of synthetic input: table clear + build + count + scan + result-size readback +
write. The headline (`value`) is BASELINE.json config 2 (16M x 256M, i32, unique build keys, 100 % match);
with N > 1 the probe relation is sharded (weak scaling: 256M probe rows per GPU, build side broadcast from rank 0 inside the step).
Next to it, under `extras`, the same invocation measures what the headline does not exercise: config 2 with sparse keys (`c2_sparse`:
no dense key range, so only the hash-table paths apply), the hash layout forced on config 2, config 3 (broadcast plan, strong
scaling), config 4, config 5 (radix join; with N > 1 the radix-partitioned plan with the exchange fused into the partition kernel,
2^28 rows per GPU and side) and the two shapes the reference published numbers for (join-performances.md) — each with its own
parity guard and roofline fraction.

`value` is device-timed (CUDA events on the launching stream, max over ranks) with inputs resident in HBM; `e2e` is the same metric
through the C-ABI call that takes HOST buffers (hjJoinHost: H2D of both relations and D2H of the pairs inside the timed region).
`--impl reference` times the reference's algorithm (the oracle's join_v1 loop restatement, OpenMP over all host cores — the reference
has no CPU lowering of its own) on the SAME config at full size.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "join tuples/sec (build+probe)"
UNIT = "tuples/s"
# the reference's own published numbers (join-performances.md:3-11,16-24, unstated hardware): seconds for "all kernels"
PUBLISHED_S = {"ref10m": {"join_v1": 2.0, "join_v2": 1.5}, "ref100m": {"join_v1": 12.0, "join_v2": 12.5}}


# ----------------------------------------------------------------------------------------------------------
def algorithmic_bytes(nR, nS, out, key_bytes, table_in_hbm=True, lookups_in_write=False, candidates=None):
    """inputs read once + one slot write per build row + one slot read per match candidate (when the table is
    HBM-resident) + outputs written once. Returned per phase so the dominant kernel gets its own share: the slot reads belong to the
    pass that does the lookups (the count pass, or the write pass when the count runs by range test)."""
    slot = 8 if key_bytes == 4 else 16
    build = nR * (key_bytes + 4) + nR * slot
    lookups = (out if candidates is None else candidates) * slot if table_in_hbm else 0
    count = nS * key_bytes + (0 if lookups_in_write else lookups)
    write = out * 8 + (lookups if lookups_in_write else 0)
    return {"build": build, "count": count, "write": write, "total": build + count + write}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 20 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(", ") for r in Path(self.f.name).read_text().strip().splitlines() if r.count(",") >= 6]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[0]) for r in rows]
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), start=3):
            if any(r[i].strip() == "Active" for r in rows):
                reasons.append(name)
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def measured_peak_gbs() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def bind_near_gpu(torch, dev) -> str:
    """Run this rank (and first-touch its pinned host buffers) on the CPUs NVML lists as local to its GPU: with eight ranks on a two-socket
    host the staging buffers otherwise land on whichever NUMA node the launcher happened to start the process on, and half the ranks
    cross the socket interconnect for every PCIe transfer (round 1: hjJoinHost 46 -> 214 ms per step from 1 to 8 ranks)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(dev)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {c for c in range(ncpu) if (mask[c // 64] >> (c % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        use = near & allowed
        if not use:
            return f"no overlap between the GPU's CPUs and the {len(allowed)} this process may use: unchanged"
        os.sched_setaffinity(0, use)
        note = f"{len(use)} CPUs local to the GPU (of {len(allowed)} allowed)"
        # pinned staging memory on the GPU's own NUMA node even when the container's CPU set spans one socket only
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node_file = Path(f"/sys/bus/pci/devices/{bdf}/numa_node")
        if node_file.exists():
            node = int(node_file.read_text().strip())
            if node >= 0:
                import ctypes
                mask = ctypes.c_ulong(1 << node)
                rc = ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), 64)      # set_mempolicy(MPOL_PREFERRED, {node})
                note += f"; memory policy: prefer NUMA node {node}" + ("" if rc == 0 else f" (refused, errno {ctypes.get_errno()})")
        return note
    except Exception as e:                                    # noqa: BLE001  (NVML missing / restricted: the run goes on unbound)
        return f"unbound ({type(e).__name__})"


def host_threads() -> int:
    """Every host core this process may run on (torchrun exports OMP_NUM_THREADS=1: the CPU arm asks explicitly)."""
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ----------------------------------------------------------------------------------------------------------
def cpu_join(cfg, reps: int, budget_s: float, probe_rows: int | None = None) -> dict:
    """The reference's algorithm on the host cores: oracle/oracle_join.c join_v1 restatement (chained table, hash = key % H,
    count -> scan -> write), OpenMP, on the FULL workload (probe_rows = None) or on its first probe_rows probe rows."""
    from oracle import Oracle
    o = Oracle()
    b, p = cfg.build, cfg.probe
    nS = p.n if probe_rows is None else min(p.n, probe_rows)
    R = o.generate(b.n, b.key_bytes, b.kind, b.seed, b.lo, b.domain, b.p16, b.key_mul)
    S = o.generate(nS, p.key_bytes, p.kind, p.seed, p.lo, p.domain, p.p16, p.key_mul)
    H = max(1, min(b.n, 2**31 - 1))          # one bucket per build row: the strongest setting of the reference's H
    threads = host_threads()
    secs, n_out, t_start = [], 0, time.perf_counter()
    for _ in range(max(1, reps)):
        n_out, sec = o.join_timed(R, S, H=H, threads=threads)
        secs.append(sec)
        if time.perf_counter() - t_start > budget_s:
            break
    what = f"all {nS} probe rows" if nS == p.n else f"first {nS} of {p.n} probe rows"
    return {"seconds": secs, "cores": threads, "n_out": n_out, "nS": nS,
            "sample": f"full build side ({b.n} rows) x {what} of {cfg.name}, H={H} buckets, {n_out} pairs, best {min(secs):.3f} s, {len(secs)} run(s)"}


def workload_name(cfg, world: int, plan: str) -> str:
    b, p = cfg.build, cfg.probe
    kt = "i32" if b.key_bytes == 4 else "i64"
    return f"{cfg.name}: {b.n} build x {p.n} probe rows, {kt} keys, {plan}" + (f", {world} GPUs" if world > 1 else "")


def run_reference_arm(args, cfg) -> None:
    if int(os.environ.get("RANK", "0")) != 0:
        return                                     # under torchrun rank 0 alone runs the CPU arm; the others exit 0 without work
    res = cpu_join(cfg, args.warmup + args.steps, budget_s=args.cpu_budget_s, probe_rows=(1 << args.cpu_sample_log2) if args.cpu_sample_log2 else None)
    timed = res["seconds"][min(args.warmup, len(res["seconds"]) - 1):]
    ms = 1e3 * sum(timed) / len(timed)
    value = (cfg.build.n + res["nS"]) / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(timed), "warmup": len(res["seconds"]) - len(timed),
            "steps_requested": args.steps, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32" if cfg.build.key_bytes == 4 else "int64", "data": "synthetic (seeded generators, same as the GPU arm)",
            "config": {"workload": workload_name(cfg, 1, "single GPU"), "build_rows": cfg.build.n, "probe_rows": res["nS"], "result_pairs": res["n_out"],
                       "note": f"CPU port of join_v1 on the host cores; stops after {args.cpu_budget_s:.0f} s of steps"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ----------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


LAYOUT_NAMES = {0: "bucketised hash", 1: "direct-address", 2: "grouped", 3: "radix-partitioned"}


def launches_per_step(layout: int, by_range: bool, nR: int, nS: int, kb: int, sparse: int, dense_policy: int) -> int:
    """Kernel launches of one step, mirroring hj_kernels.cu / hj_radix.cu (memsets and copies are not kernels)."""
    n = 2 + (1 if nR else 0) + (1 if nR >= (1 << 18) else 0)                # k_init_header, k_decide, k_minmax, k_sample_dups
    scan = lambda m: 2 if m > 16384 else 1                                   # noqa: E731  k_scan_blocks (+ k_scan_add)
    if layout == 3:
        parts = 1 << max(2, min(16, (max(nR, 1) - 1).bit_length() - 12))
        n += 2 * 4 + 1                                                        # two partition passes (blocks, hist, scan, scatter) + k_rj_header
        n += 2 * 4 + 1 + scan(parts) + 1 + 1 + scan(nS // 16384 + 65537) + 1  # probe side: partition, item counts, scan, items, count join, scan; write join
        return n
    n += 1 + (4 if dense_policy else 0) + 6                                   # k_clear, [dense, dense_verify, fallback_prepare, clear], hash, group_prepare, clear, group count/offsets/fill
    chunks = -(-nS // (16384 if kb == 4 else 1024))
    n += (1 if sparse and nS >= (1 << 20) and layout != 2 else 0) + 1 + (1 if sparse and layout != 2 else 0) + scan(chunks) + 1
    return n


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c2s", "c3", "c4", "c5", "ref10m", "ref100m"])
    ap.add_argument("--scale-log2", type=int, default=0, help="shrink (<0) the workload by 2^k rows on both sides (debug only)")
    ap.add_argument("--cpu-sample-log2", type=int, default=0, help="CPU arm: only the first 2^k probe rows (0 = the full workload, the default)")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="CPU arm: stop starting new steps after this many seconds")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload (no c2_sparse / hash / c3 / c4 / c5 / ref shapes)")
    ap.add_argument("--extras", default="", help="comma list restricting the extras (names: c2_sparse,hash_layout,match_cache,fused,c3,c4,c5,ref10m,ref100m)")
    ap.add_argument("--c5-total-log2", type=int, default=0, help="c5 only: fix the TOTAL rows per side at 2^k (strong scaling); default 2^28 rows per GPU (weak)")
    ap.add_argument("--overlap-build", action="store_true", help="c5, N > 1, fused exchange: local build on a second stream while the probe side is pushed")
    ap.add_argument("--exchange", default="staged", choices=["fused", "staged", "nccl"], help="c5, N > 1: parts staged locally and carried by the copy engines (default), peer-store partition kernel, or partition + NCCL all-to-all")
    ap.add_argument("--layout", default="auto", choices=["auto", "cache", "hash"],
                    help="auto = library default (direct-address table for dense key ranges, counted by range test when gap-free and unique); "
                         "cache = direct-address table with the match cache only; hash = never the direct-address layout")
    ap.add_argument("--sparse", type=int, default=1, choices=[0, 1, 2], help="hit lists for selective joins (0 never, 1 sampled on the device, 2 always)")
    ap.add_argument("--partition-threads", type=int, default=0, choices=[0, 256, 512, 1024], help="experiment: CTA shape of the partition scatter kernel")
    ap.add_argument("--no-sliced", action="store_true", help="tables of 48 MB .. 1 GB of buckets: radix layout instead of one slice-ordered hash table")
    ap.add_argument("--no-radix", action="store_true", help="tables beyond L2 reach: one hash table in global memory instead of the radix join")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # Rank 0 prints ONE JSON line on stdout and nothing else does: libraries that write to fd 1 (NCCL prints its version banner there
    # when a site-wide nccl.conf sets NCCL_DEBUG=VERSION) are sent to stderr for the whole run; _emit() writes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    import __graft_entry__ as g
    from mlir_hashjoin_b200 import datagen

    def make_config(name: str, world: int):
        cfg = datagen.config(name.upper(), args.scale_log2)
        if name == "c5":
            per_gpu = 1 << 28                       # rows per GPU per side: 2e9-row C5 at 8 GPUs is 2.5e8 rows per GPU
            n = per_gpu * world if args.scale_log2 == 0 else max(1, (per_gpu * world) >> -args.scale_log2)
            if args.c5_total_log2:
                n = 1 << args.c5_total_log2
            cfg = datagen.JoinConfig("C5", datagen.replace(cfg.build, n=n, domain=n), datagen.replace(cfg.probe, n=n, domain=n), n, cfg.note)
        return cfg

    if args.impl == "reference":
        from oracle import build_oracle
        build_oracle()
        run_reference_arm(args, make_config(args.workload, 1))
        return

    g.build()
    import torch
    import torch.distributed as dist
    from mlir_hashjoin_b200 import _lib, join
    from mlir_hashjoin_b200 import dist as hjdist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the join has no CPU fallback (use --impl reference for the CPU arm)")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # keeps NCCL's banner off stdout: rank 0 prints ONE JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    stream = torch.cuda.current_stream()
    peak, peak_src = measured_peak_gbs()
    DENSE_POLICY = {"hash": 0, "cache": 1, "auto": 2}
    lib.hjSetSparse(args.sparse)
    lib.hjSetLocality(0 if args.no_radix else 1)
    lib.hjSetSliced(0 if args.no_sliced else 1)
    if args.partition_threads:
        lib.hjSetPartitionThreads(args.partition_threads)

    def allmax(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def allsum_int(x: int) -> int:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def all_digest(d: tuple[int, int]) -> tuple[int, int]:
        """(sum mod 2^64, xor) of the ranks' order-independent pair digests."""
        if world == 1:
            return d
        signed = [v - (1 << 64) if v >= (1 << 63) else v for v in d]
        mine = torch.tensor(signed, dtype=torch.int64, device=dev)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        s_, x_ = 0, 0
        for t in every:
            a_, b_ = [int(v) & ((1 << 64) - 1) for v in t.tolist()]
            s_ = (s_ + a_) & ((1 << 64) - 1); x_ ^= b_
        return s_, x_

    # ------------------------------------------------------------------------------------------------------
    # one workload: inputs resident in HBM, W warm-up steps, K timed steps, parity guard on the last step
    # plan: "single" | "broadcast" (probe rows sharded, build side broadcast in the step) | "radix" (both sides sharded, exchange + local join)
    # ------------------------------------------------------------------------------------------------------
    def measure(name: str, steps: int, warmup: int, layout: str = "auto", sample_clocks: bool = False, strong: bool = False, fused: bool = False) -> dict:
        cfg = make_config(name, world)
        b, p = cfg.build, cfg.probe
        kb = b.key_bytes
        plan = "single" if world == 1 else ("radix" if name == "c5" else "broadcast")
        lib.hjSetAllowDense(DENSE_POLICY[layout])
        if plan == "radix":
            blo, bhi = hjdist.shard_range(b.n, rank, world)
            plo, phi = hjdist.shard_range(p.n, rank, world)
            dR = datagen.generate(b, dev, blo, bhi - blo)
            dS = datagen.generate(p, dev, plo, phi - plo)
            nR_job, nS_job, nS_rank = b.n, p.n, phi - plo
            peer_x = (hjdist.PeerExchange(int(1.25 * b.n / world) + 65536, b.dtype, dev), hjdist.PeerExchange(int(1.25 * p.n / world) + 65536, p.dtype, dev)) if args.exchange in ("fused", "staged") else None
            table = join.allocateHashTable(int(1.25 * b.n / world) + 65536, None, b.dtype, dev)
        else:
            if strong:                                          # the config's probe relation split over the ranks (config 3: "1M x 1B at 1/2/4/8 GPUs")
                plo, phi = hjdist.shard_range(p.n, rank, world)
                dS = datagen.generate(p, dev, plo, phi - plo)
                nS_job = p.n
            else:                                               # weak: every rank probes its own p.n rows (global probe relation = world * p.n rows)
                plo = rank * p.n
                dS = datagen.generate(datagen.replace(p, n=p.n * world), dev, plo, p.n)
                nS_job = p.n * world
            nS_rank = dS.numel()
            dR = datagen.generate(b, dev) if rank == 0 or world == 1 else torch.empty(b.n, dtype=b.dtype, device=dev)
            nR_job = b.n
            peer_x = None
            table = join.allocateHashTable(b.n, None, b.dtype, dev)
        out = {"R": None, "S": None}             # result columns live outside the timed region (allocation is excluded)

        def result_columns(n):
            if out["R"] is None or out["R"].numel() < n:
                out["R"] = out["S"] = None
                out["R"] = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
                out["S"] = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            return out["R"][:n], out["S"][:n]
        torch.cuda.synchronize()
        n_out = [0]
        marks, leg_events = {}, []

        def step():
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if plan == "radix":
                ev[0].record(stream)
                ev[1].record(stream); ev[2].record(stream)              # moved by the `exchanged` hook: partition + exchange | local join
                if peer_x:
                    marks.clear()
                    hook = lambda: (ev[1].record(stream), ev[2].record(stream))      # noqa: E731
                    if args.exchange == "staged":
                        a, bb = hjdist.radix_join_staged(dR, blo, dS, plo, *peer_x, exchanged=hook, table=table, marks=marks, result=result_columns)
                    else:
                        a, bb = hjdist.radix_join_fused(dR, blo, dS, plo, *peer_x, exchanged=hook, table=table, marks=marks, overlap_build=args.overlap_build,
                                                        result=result_columns)
                    leg_events.append(dict(marks))
                else:
                    a, bb = hjdist.radix_join(dR, blo, dS, plo, table=table)
                ev[3].record(stream)
                n_out[0] = a.numel()
                return ev, (a, bb)
            ev[0].record(stream)
            if world > 1:
                dist.broadcast(dR, src=0)                       # the build side travels once per step (NCCL over NVLink)
            join.initializeHashTable(table)
            join.buildTable(dR, table)
            ev[1].record(stream)
            if fused:                                           # single-pass probe into a result the caller bounded (|S| pairs for a unique build)
                oR, oS = result_columns(nS_rank)
                n = join.join_fused(dS, table, oR, oS, probeRowBase=plo)
                ev[2].record(stream); ev[3].record(stream)
                n_out[0] = n
                return ev, (oR[:n], oS[:n])
            n = join.countRows(dS, table, probeRowBase=plo)      # includes the result-size readback (host sync); the probe row ids are stated here (hjCountRows)
            ev[2].record(stream)
            oR, oS = result_columns(n)
            if n:
                join.probeRelation(dS, table, oR, oS, probeRowBase=plo)
            ev[3].record(stream)
            n_out[0] = n
            return ev, (oR, oS)

        sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None   # nvidia-smi needs ~100 ms to start: launch it before the warm-up
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        all_ev, last = [], None
        for _ in range(steps):
            ev, last = step()
            all_ev.append(ev)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3 / steps
        clocks = sampler.stop() if sampler else None
        ms = allmax(sum(e[0].elapsed_time(e[3]) for e in all_ev) / steps)
        wall_ms = allmax(wall_ms)
        mean = lambda i, j: sum(e[i].elapsed_time(e[j]) for e in all_ev) / steps     # noqa: E731
        phases = {"partition_exchange": mean(0, 1), "local_join": mean(2, 3)} if plan == "radix" else {"build": mean(0, 1), "count": mean(1, 2), "write": mean(2, 3)}
        tot_out = allsum_int(n_out[0])
        layout_code = lib.hjTableLayout(table.storage.data_ptr(), None)
        hit_lists = bool(plan != "radix" and not fused and table.scratch is not None and (layout_code & 0xFF) in (0, 1) and lib.hjProbePath(table.scratch.data_ptr(), nS_rank, kb, None) == 1)

        # ---- parity guard on the last timed step: count + size-independent properties, all on the device ---------------------
        outR, outS = last
        parity = {}
        if cfg.expected_out is not None:
            expect = cfg.expected_out * (world if plan == "broadcast" and not strong else 1)
            parity["count_matches_analytic"] = tot_out == expect
        if plan == "radix":
            kR = datagen.generate_at(b, outR); kS = datagen.generate_at(p, outS)             # rows live on other ranks: regenerate their keys
            parity["all_pairs_join_equal_keys"] = bool(allsum_int(int((kR != kS).sum().item())) == 0)
            want = join.pair_digest(torch.zeros(nS_rank, dtype=torch.int32, device=dev), torch.arange(plo, plo + nS_rank, dtype=torch.int64, device=dev).to(torch.int32))
            got = join.pair_digest(torch.zeros_like(outS), outS)
            parity["probe_rows_are_a_permutation"] = all_digest(got) == all_digest(want)     # over all ranks: every probe row of [0, n) exactly once
            del kR, kS
        elif outR.numel():
            local = (outS.long() - plo)
            if rank == 0 or world == 1:
                parity["all_pairs_join_equal_keys"] = bool((dR[outR.long()] == dS[local]).all())
            if cfg.expected_out is not None and name in ("c2", "c2s", "c5"):                 # unique build, every probe row matches once
                parity["probe_rows_are_a_permutation"] = join.pair_digest(torch.zeros_like(outS), outS) == \
                    join.pair_digest(torch.zeros(nS_rank, dtype=torch.int32, device=dev), torch.arange(plo, plo + nS_rank, dtype=torch.int64, device=dev).to(torch.int32))
            elif cfg.expected_out is None and (rank == 0 or world == 1):                     # exact size, computed without the join kernels
                sR, _ = torch.sort(dR)
                parity["count_matches_sorted_search"] = int((torch.searchsorted(sR, dS, right=True) - torch.searchsorted(sR, dS, right=False)).sum().item()) == n_out[0]
                del sR
            del local
        res = {"workload": workload_name(cfg, world, {"single": "single GPU", "broadcast": "broadcast build, probe sharded" + (" (strong)" if strong else " (weak)"),
                                                     "radix": "radix partition, " + {"fused": "exchange fused into the partition kernel (NVLink peer stores)", "nccl": "NCCL all-to-all",
                                                                                    "staged": "parts staged locally and carried by the copy engines beside the partition / build kernels"}[args.exchange]}[plan]),
               "ms_per_step": ms, "wall_ms_per_step": wall_ms, "value": (nR_job + nS_job) / (ms / 1e3), "unit": UNIT, "phases_ms": phases,
               "build_rows": nR_job, "probe_rows": nS_job, "result_pairs": tot_out, "pairs_per_s": tot_out / (ms / 1e3), "parity": parity,
               "table_layout_chosen": LAYOUT_NAMES.get(layout_code & 0xFF, "?") + (", slice-ordered" if layout_code & 0x200 else "") + (", count by range test" if layout_code & 0x100 and not hit_lists else "") + (", hit lists" if hit_lists else ""),
               "steps": steps, "clocks": clocks}
        # ---- roofline: algorithmic bytes over the event-timed phases -----------------------------------------------
        if plan == "radix":
            sent = (world - 1) / world * (dR.numel() + dS.numel()) * (kb + 4)            # (key, global row id) tuples leaving this GPU
            px, lj = phases["partition_exchange"], phases["local_join"]
            n_loc = b.n // world
            res["c5_phases"] = {"partition_exchange_ms": px, "local_join_ms": lj, "nvlink_bytes_sent_per_gpu": int(sent), "nvlink_achieved_gbs": sent / (px / 1e3) / 1e9,
                                "nvlink_peak_gbs": 770.0, "nvlink_frac": sent / (px / 1e3) / 1e9 / 770.0,
                                "local_join_hbm_frac": 64 * n_loc / (lj / 1e3) / 1e9 / peak,
                                "note": "nvlink peak = measured peer copy per direction per GPU (B200_PROFILING.md); the exchange phase also holds the owner histograms of both relations, the all-gather of the count matrices and its one host read; local join bytes = 64 per row (SURVEY 8d, C5)"}
            res["gpu_launches_per_step"] = 2 * 4 + launches_per_step(3, False, n_loc, n_loc, kb, args.sparse, DENSE_POLICY[layout])
            if leg_events:                                          # legs of the exchange phase on this rank, mean over the timed steps
                legs = leg_events[-steps:]
                mean = lambda a_, b_: sum(l[a_].elapsed_time(l[b_]) for l in legs) / len(legs)      # noqa: E731
                if os.environ.get("HJ_BENCH_DEBUG_LEGS"):
                    for l in legs:
                        ks = list(l)
                        print("legs", {k: round(l["start"].elapsed_time(l[k]), 2) for k in ks}, file=sys.stderr)
                if args.exchange == "staged":
                    names = ["start", "histogram_build", "count_matrix_build", "scatter_build", "histogram_probe", "count_matrix_probe", "scatter_probe",
                             "local_build", "local_count", "local_write"]   # the SM stream
                    res["c5_phases"]["exchange_legs_ms"] = {b_: mean(a_, b_) for a_, b_ in zip(names, names[1:])}
                    ce_ms = {"build": mean("copy_build_start", "copy_build_end"), "probe": mean("copy_probe_start", "copy_probe_end")}
                    res["c5_phases"]["copy_engine_ms"] = ce_ms      # second stream, beside the legs above: copies + the landing barrier
                    res["c5_phases"]["exchange_done_ms"] = mean("start", "copy_probe_end")
                    gbs = sent / ((ce_ms["build"] + ce_ms["probe"]) / 1e3) / 1e9
                    res["c5_phases"].update({"nvlink_achieved_gbs": gbs, "nvlink_frac": gbs / 770.0,
                                             "note": "staged plan: partition_exchange_ms runs until the probe side has landed and INCLUDES the local build that runs beside the probe side's copies; "
                                                     "nvlink figures = bytes sent / time of the copy legs on the second stream; local join bytes = 64 per row (SURVEY 8d, C5)"})
                else:
                    names = ["start", "histograms", "count_matrix", "push_build", "push_probe", "local_build", "local_count", "local_write"]
                    res["c5_phases"]["exchange_legs_ms"] = {b_: mean(a_, b_) for a_, b_ in zip(names, names[1:])}
                    res["c5_phases"]["overlap_build"] = bool(args.overlap_build)
        else:
            by_range = bool(layout_code & 0x100) and not hit_lists
            table_in_hbm = lib.hjTableBytes(b.n, kb) > 96 * 2**20
            ab = algorithmic_bytes(b.n, nS_rank, n_out[0], kb, table_in_hbm=table_in_hbm, lookups_in_write=by_range)
            dom = max(phases, key=phases.get)
            kernel = {"build": "build sequence (k_minmax, k_clear, k_build_dense | k_build_hash | k_group_* | k_rp_hist + k_rp_scatter x 2)",
                      "count": "k_count | k_count_sparse | k_count_range | k_rp_* x 2 + k_rj_join<count> (+ k_sample_hits, k_scan_blocks and the result-size readback)",
                      "write": "k_write | k_write_sparse | k_write_range | k_rj_join<write>"}[dom]
            ach = ab[dom] / (phases[dom] / 1e3) / 1e9
            res["roofline"] = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                               "algorithmic_bytes": ab[dom], "kernel_ms": phases[dom], "peak_source": peak_src, "phases_ms": phases,
                               "share_of_step": phases[dom] / (sum(phases.values()) or 1),
                               "job": {"algorithmic_bytes": ab["total"], "achieved": ab["total"] / (ms / 1e3) / 1e9, "frac": ab["total"] / (ms / 1e3) / 1e9 / peak}}
            res["gpu_launches_per_step"] = launches_per_step(layout_code & 0xFF, by_range, b.n, nS_rank, kb, args.sparse, DENSE_POLICY[layout])
        res["_keep"] = (cfg, dR, dS, plo, n_out[0])
        del out, last, outR, outS, table
        return res

    # ---- headline --------------------------------------------------------------------------------------------------------------
    head = measure(args.workload, args.steps, args.warmup, layout=args.layout, sample_clocks=True, strong=(args.workload == "c3" and world > 1))
    cfg, dR, dS, plo, n_out_rank = head.pop("_keep")
    b, p = cfg.build, cfg.probe
    kb = b.key_bytes

    # ---- end to end through the host-buffer C-ABI call (rank-local; H2D + D2H inside the timed region) --------------------------
    e2e = None
    if not args.no_e2e and args.workload != "c5":
        affinity = bind_near_gpu(torch, dev) if world > 1 else "single rank: unbound"
        hR = torch.empty(b.n, dtype=b.dtype, pin_memory=True); hS = torch.empty(dS.numel(), dtype=p.dtype, pin_memory=True)
        if world > 1:
            dist.broadcast(dR, src=0)
        hR.copy_(dR); hS.copy_(dS)
        cap = n_out_rank
        hOr = torch.empty(max(cap, 1), dtype=torch.int32, pin_memory=True); hOs = torch.empty(max(cap, 1), dtype=torch.int32, pin_memory=True)
        torch.cuda.empty_cache()
        times = []
        for i in range(1 + args.e2e_steps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t1 = time.perf_counter()
            got = lib.hjJoinHost(hR.data_ptr(), b.n, hS.data_ptr(), dS.numel(), kb, hOr.data_ptr(), hOs.data_ptr(), cap)
            dt = time.perf_counter() - t1
            if got != cap:
                raise SystemExit(f"hjJoinHost returned {got}, expected {cap}: {lib.hjLastErrorString()}")
            if i > 0:
                times.append(dt)
        te = allmax(sum(times) / len(times))
        h2d, d2h = (b.n + dS.numel()) * kb, cap * 8
        e2e = {"value": (head["build_rows"] + head["probe_rows"]) / te, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": te * 1e3, "pcie_gbs_per_rank": {"h2d": h2d / te / 1e9, "d2h": d2h / te / 1e9}, "host_affinity": affinity,
               "api": "hjJoinHost (include/hashjoin_b200.h), pinned host buffers; probe relation streamed through 3 device chunks, pairs through 2 result slots"}
        lib.hashJoinRelease()
        del hR, hS, hOr, hOs
    del dR, dS
    torch.cuda.empty_cache()

    # ---- extras: what the headline does not exercise, each with its own parity guard and roofline -------------------------------
    extras = {}
    if not args.no_extras and args.workload == "c2" and args.layout == "auto":
        want = set(x for x in args.extras.split(",") if x)
        ex_steps = max(5, min(20, args.steps // 5))

        def extra(key, name, **kw):
            if want and key not in want:
                return
            r = measure(name, ex_steps, 3, **kw)
            r.pop("_keep", None)
            extras[key] = r
            torch.cuda.empty_cache()
        extra("c2_sparse", "c2s")                                # same sizes, no dense key range: the hash-table path the north star names
        extra("hash_layout", "c2", layout="hash")                # config 2 itself with the direct-address layout switched off
        extra("match_cache", "c2", layout="cache")               # direct-address table, lookups in the count pass (match cache)
        extra("fused", "c2", fused=True)                         # build + ONE probe pass (hjJoinFused) into a result bounded by |S|
        extra("c3", "c3", strong=world > 1)
        extra("c4", "c4")
        extra("c5", "c5")
        if world == 1:
            for key in ("ref10m", "ref100m"):
                extra(key, key)
                if key in extras:
                    pub = PUBLISHED_S[key]
                    n_total = extras[key]["build_rows"] + extras[key]["probe_rows"]
                    extras[key]["reference_published"] = {"source": "join-performances.md:3-11,16-24 (hardware unstated, sm_86 cubin)", "seconds": pub,
                                                          "tuples_per_s": {k: n_total / v for k, v in pub.items()},
                                                          "speedup_vs_published": {k: v * 1e3 / extras[key]["ms_per_step"] for k, v in pub.items()}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:              # a reported baseline, timed on rank 0 at N = 1 only, on the FULL workload
        r = cpu_join(cfg, 1, budget_s=0.0, probe_rows=(1 << args.cpu_sample_log2) if args.cpu_sample_log2 else None)
        cpu = {"value": (b.n + r["nS"]) / min(r["seconds"]), "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    roofline = head.get("roofline")
    traffic_file = ROOT / "profiles" / "traffic.json"
    if roofline is not None and traffic_file.exists():
        import hashlib
        entry = json.loads(traffic_file.read_text()).get(args.workload, {}).get(head["table_layout_chosen"])
        if entry:        # per-launch DRAM bytes of the dominant kernel from one ncu --set full capture (tools/make_traffic.py), valid for the source it was taken on
            src = ROOT / entry["source_file"]
            fresh = src.exists() and hashlib.sha256(src.read_bytes()).hexdigest() == entry["source_sha256"]
            roofline["traffic"] = entry["dram_bytes"] if fresh else None
            roofline["traffic_source"] = f'{entry["kernel"]}: {entry["source"]}' + ("" if fresh else " — STALE: the kernel source changed since the capture, figure withheld")

    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong" if (args.workload == "c5" and args.c5_total_log2) or (args.workload == "c3" and world > 1) else "weak",
            "vs_baseline": None, "dtype": "int32" if kb == 4 else "int64", "data": "synthetic (seeded device generators, bit-identical to the oracle's)",
            "config": {"workload": head["workload"], "build_rows": head["build_rows"], "probe_rows": head["probe_rows"], "result_pairs": head["result_pairs"],
                       "l2_hygiene": "inputs larger than L2 (probe column >= 1 GiB per GPU streams through every step)",
                       "timing": "CUDA events on the launching stream per step, summed over steps, max over ranks",
                       "wall_ms_per_step": head["wall_ms_per_step"], "table_layout": args.layout, "table_layout_chosen": head["table_layout_chosen"]},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": head["gpu_launches_per_step"] * args.steps, "clocks": head["clocks"], "parity": head["parity"]}
    if "c5_phases" in head:
        line["c5_phases"] = head["c5_phases"]
    if args.workload in PUBLISHED_S:
        line["vs_baseline"] = PUBLISHED_S[args.workload]["join_v2"] * 1e3 / head["ms_per_step"]
    for key, r in extras.items():                              # flat keys the driver keeps: <name> = compact result of that arm
        line[key] = {k: v for k, v in r.items() if k not in ("clocks",)}
    if "hash_layout" in extras:
        line["hash_layout"]["note"] = "hjSetAllowDense(0): same inputs as the headline, direct-address layout disabled (the hash-table path on config 2 itself)"
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
