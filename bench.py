#!/usr/bin/env python
"""bench.py — join tuples/s (build + probe) for the B200 hash join, one JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

A step is one pass of the hot path over one batch of synthetic input: table clear + build + count + scan +
result-size readback + write. N=1 runs BASELINE.json config 2 (16M x 256M, i32, unique build
keys, 100 % match). N>1 shards the probe relation (weak scaling: 256M probe rows per GPU, build side broadcast from
rank 0 inside the step), or `--workload c5` radix-partitions both sides and shuffles them with an NCCL all-to-all.

`value` is device-timed (CUDA events on the launching stream, max over ranks) with inputs resident in HBM; `e2e` is the
same metric through the C-ABI call that takes HOST buffers (hjJoinHost: H2D of both relations and D2H of the pairs
inside the timed region). `--impl reference` times the reference's algorithm (the oracle's join_v1 loop restatement,
OpenMP over all host cores — the reference has no CPU lowering of its own) on a bounded sample.
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


# ----------------------------------------------------------------------------------------------------------
def algorithmic_bytes(nR, nS, out, key_bytes, matches_per_hit=1.0, table_in_hbm=True, lookups_in_write=False):
    """inputs read once + one slot write per build row + one slot read per match candidate (when the
    table is HBM-resident) + outputs written once. Returned per kernel so the dominant kernel gets its own share: the slot reads
    belong to the pass that does the lookups (the count pass, or the write pass when the count runs by range test)."""
    slot = 8 if key_bytes == 4 else 16
    build = nR * (key_bytes + 4) + nR * slot
    lookups = out * slot if table_in_hbm else 0
    count = nS * key_bytes + (0 if lookups_in_write else lookups)
    write = out * 8 + (lookups if lookups_in_write else 0)
    return {"build": build, "count": count, "write": write, "total": build + count + write}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 100 ms while the timed region runs."""
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


# ----------------------------------------------------------------------------------------------------------
def cpu_baseline(cfg, sample_probe_log2: int, reps: int = 1, threads: int = 0) -> dict:
    """The reference's algorithm on the host cores: oracle/oracle_join.c join_v1 restatement (chained table,
    hash = key % H, count -> scan -> write), OpenMP. Bounded sample: the FULL build side, the first 2^k probe rows."""
    from oracle import Oracle
    o = Oracle()
    b, p = cfg.build, cfg.probe
    nS = min(p.n, 1 << sample_probe_log2)
    R = o.generate(b.n, b.key_bytes, b.kind, b.seed, b.lo, b.domain, b.p16, b.key_mul)
    S = o.generate(nS, p.key_bytes, p.kind, p.seed, p.lo, p.domain, p.p16, p.key_mul)
    H = max(1, min(b.n, 2**31 - 1))          # one bucket per build row: the strongest setting of the reference's H
    if threads == 0:                          # every host core this process may run on (torchrun exports OMP_NUM_THREADS=1: ask explicitly)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    secs, n_out = [], 0
    for _ in range(reps):
        n_out, sec = o.join_timed(R, S, H=H, threads=threads)
        secs.append(sec)
    best = min(secs)
    cores = threads
    return {"value": (b.n + nS) / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"full build side ({b.n} rows) x first {nS} probe rows of {cfg.name}, H={H} buckets, {n_out} pairs, {best:.3f} s",
            "seconds": secs}


def run_reference_arm(args, cfg) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_baseline(cfg, args.cpu_sample_log2, reps=args.warmup + args.steps)
    t_all = res["seconds"][args.warmup:]
    nS = min(cfg.probe.n, 1 << args.cpu_sample_log2)
    ms = 1e3 * sum(t_all) / len(t_all)
    value = (cfg.build.n + nS) / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32" if cfg.build.key_bytes == 4 else "int64",
            "data": "synthetic (seeded generators, same as the GPU arm)",
            "config": {"workload": workload_name(args, cfg), "sample": res["sample"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def workload_name(args, cfg) -> str:
    b, p = cfg.build, cfg.probe
    kt = "i32" if b.key_bytes == 4 else "i64"
    if args.workload == "c5" and args.gpus > 1:
        how = "partition kernel stores into peer receive buffers over NVLink" if args.exchange == "fused" else "partition + NCCL all-to-all"
        return f"{cfg.name}: {b.n} build x {p.n} probe rows in total over {args.gpus} GPUs, {kt} keys, radix partition ({how})"
    plan = "single GPU" if args.gpus == 1 else "broadcast build, probe sharded"
    return f"{cfg.name}: {b.n} build x {p.n} probe rows per GPU, {kt} keys, {plan}"


# ----------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scale-log2", type=int, default=0, help="shrink (<0) the workload by 2^k rows on both sides (debug only)")
    ap.add_argument("--cpu-sample-log2", type=int, default=26, help="probe rows of the CPU baseline sample (2^k)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--c5-total-log2", type=int, default=0, help="c5 only: fix the TOTAL rows per side at 2^k (strong scaling); default 2^28 rows per GPU (weak)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"], help="c5 only: peer-store partition kernel vs partition + NCCL all-to-all")
    ap.add_argument("--layout", default="auto", choices=["auto", "cache", "hash"],
                    help="auto = library default (hjSetAllowDense(2): direct-address table for dense key ranges, counted by range test when gap-free and unique); "
                         "cache = direct-address table with the match cache only (hjSetAllowDense(1)); hash = force the bucketised hash table (hjSetAllowDense(0))")
    ap.add_argument("--sparse", type=int, default=1, choices=[0, 1, 2], help="hjSetSparse: hit lists for selective joins (0 never, 1 sampled on the device, 2 always)")
    ap.add_argument("--dense-waves", type=int, default=None, help="hjSetDenseWaves (experiment): grid of the direct-address probe kernels")
    ap.add_argument("--no-hash-arm", action="store_true", help="skip the extra forced-hash-layout measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # Rank 0 prints ONE JSON line on stdout and nothing else does: libraries that write to fd 1 (NCCL prints its version banner there
    # when a site-wide nccl.conf sets NCCL_DEBUG=VERSION) are sent to stderr for the whole run; _emit() writes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as g
    from mlir_hashjoin_b200 import datagen
    cfg = datagen.config(args.workload.upper(), args.scale_log2)
    if args.workload == "c5":
        per_gpu = 1 << 28                       # rows per GPU per side: 2e9-row C5 at 8 GPUs is 2.5e8 rows per GPU
        n = per_gpu * args.gpus if args.scale_log2 == 0 else max(1, (per_gpu * args.gpus) >> -args.scale_log2)
        if args.c5_total_log2:
            n = 1 << args.c5_total_log2
        cfg = datagen.JoinConfig("C5", datagen.replace(cfg.build, n=n, domain=n), datagen.replace(cfg.probe, n=n, domain=n), n, cfg.note)

    if args.impl == "reference":
        from oracle import build_oracle
        build_oracle()
        run_reference_arm(args, cfg)
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
    b, p = cfg.build, cfg.probe
    kb = b.key_bytes

    # ---- inputs, resident in HBM before the timed region ------------------------------------------------------
    if args.workload == "c5" and world > 1:
        blo, bhi = hjdist.shard_range(b.n, rank, world)
        plo, phi = hjdist.shard_range(p.n, rank, world)
        dR = datagen.generate(b, dev, blo, bhi - blo)
        dS = datagen.generate(p, dev, plo, phi - plo)
        nR_job, nS_job = b.n, p.n
    else:
        # weak scaling: every rank probes its own p.n rows (global probe relation = world * p.n rows); build side replicated
        pw = datagen.replace(p, n=p.n * world)
        plo = rank * p.n
        dS = datagen.generate(pw, dev, plo, p.n)
        dR = datagen.generate(b, dev) if rank == 0 or world == 1 else torch.empty(b.n, dtype=b.dtype, device=dev)
        nR_job, nS_job = b.n, p.n * world
    peer_x = None
    if args.workload == "c5" and world > 1 and args.exchange == "fused":
        peer_x = (hjdist.PeerExchange(int(1.25 * b.n / world) + 65536, b.dtype, dev), hjdist.PeerExchange(int(1.25 * p.n / world) + 65536, p.dtype, dev))
    table = join.allocateHashTable(b.n if not (args.workload == "c5" and world > 1) else int(1.25 * b.n / world) + 1024, None, b.dtype, dev)
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    DENSE_POLICY = {"hash": 0, "cache": 1, "auto": 2}
    lib.hjSetAllowDense(DENSE_POLICY[args.layout])
    lib.hjSetSparse(args.sparse)
    if args.dense_waves is not None:
        lib.hjSetDenseWaves(args.dense_waves)
    out_buf = {"R": None, "S": None}             # result columns live outside the timed region (allocation is excluded)

    def result_columns(n):
        if out_buf["R"] is None or out_buf["R"].numel() < n:
            out_buf["R"] = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            out_buf["S"] = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        return out_buf["R"][:n], out_buf["S"][:n]
    n_out = [0]
    launches = [0]
    range_policy = [args.layout == "auto"]                     # hjSetAllowDense(2): k_count_range and k_write_range are queued too
    phase_ms = {"build": [], "count": [], "write": []}

    def step(timed: bool):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        if args.workload == "c5" and world > 1:
            ev[0].record(stream)
            ev[1].record(stream)                                    # overwritten below on the fused path: partition + exchange | local join
            a, bb = hjdist.radix_join_fused(dR, blo, dS, plo, *peer_x, exchanged=lambda: ev[1].record(stream)) if peer_x else hjdist.radix_join(dR, blo, dS, plo)
            ev[3].record(stream)
            n_out[0] = a.numel()
            launches[0] += 2 * 3 + 5
            return ev, None
        ev[0].record(stream)
        if world > 1:
            dist.broadcast(dR, src=0)                       # the build side travels once per step (NCCL over NVLink)
        join.initializeHashTable(table)
        join.buildTable(dR, table)
        ev[1].record(stream)
        n = join.countRows(dS, table)                       # includes the result-size readback (host sync)
        ev[2].record(stream)
        outR, outS = result_columns(n)
        if n:
            join.probeRelation(dS, table, outR, outS, probeRowBase=plo)
        ev[3].record(stream)
        n_out[0] = n
        # 13 build-sequence launches, k_sample_hits (>= 2^20 probe rows), 3 k_count + 2 k_count_sparse instantiations, k_scan_blocks
        # (+ k_scan_add beyond 16384 chunks), 2 k_write instantiations + k_write_sparse (the kernels whose layout / probe path was
        # not chosen exit at once; they are launches all the same)
        chunk_rows = 16384 if dS.element_size() == 4 else 1024
        launches[0] += 13 + (1 if dS.numel() >= (1 << 20) and args.sparse else 0) + 3 + (2 if args.sparse else 0) \
            + (2 if dS.numel() > 16384 * chunk_rows else 1) + 2 + (1 if args.sparse else 0) + (2 if range_policy[0] else 0)
        return ev, (outR, outS)

    def run_timed(steps, warmup, sample_clocks):
        sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None   # nvidia-smi needs ~100 ms to start: launch it before the warm-up
        for _ in range(warmup):
            step(False)
        torch.cuda.synchronize()
        launches[0] = 0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        all_ev, last = [], None
        for _ in range(steps):
            ev, last = step(True)
            all_ev.append(ev)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        clocks = sampler.stop() if sampler else None
        dev_ms = sum(e[0].elapsed_time(e[3]) for e in all_ev)
        phases = {"build": [], "count": [], "write": []}
        if args.workload == "c5" and world > 1:
            phases = {"partition_exchange": [e[0].elapsed_time(e[1]) for e in all_ev], "local_join": [e[1].elapsed_time(e[3]) for e in all_ev]}
        else:
            for e in all_ev:
                phases["build"].append(e[0].elapsed_time(e[1])); phases["count"].append(e[1].elapsed_time(e[2])); phases["write"].append(e[2].elapsed_time(e[3]))
        t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max, wall_ms_max = t.tolist()
        return dev_ms_max / steps, wall_ms_max / steps, phases, clocks, last

    ms_per_step, wall_ms_per_step, phase_ms, clocks, last = run_timed(args.steps, args.warmup, True)
    table_layout = lib.hjTableLayout(table.storage.data_ptr(), None) if not (args.workload == "c5" and world > 1) else 0   # 1 = direct-address, +0x100 = gap-free and unique
    hit_lists = bool(table.scratch is not None and table_layout and lib.hjProbePath(table.scratch.data_ptr(), dS.numel(), kb, None) == 1)
    if hit_lists:
        table_layout &= 0xFF                                  # selective join: the device-side sample chose hit lists over count-by-range
    timed_launches = launches[0]
    tot_out = torch.tensor([n_out[0]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot_out, op=dist.ReduceOp.SUM)
    value = (nR_job + nS_job) / (ms_per_step / 1e3)

    # ---- parity guard on the last timed step (cheap, device side) ---------------------------------------------
    parity = None
    if last is not None and cfg.expected_out is not None:
        outR, outS = last
        ok_count = int(tot_out.item()) == cfg.expected_out * (world if args.workload != "c5" else 1)
        idx = torch.randint(0, max(1, outR.numel()), (1 << 20,), device=dev)
        ok_keys = bool((dR[outR[idx].long()] == dS[(outS[idx] - plo).long()]).all()) if outR.numel() and (rank == 0 or world == 1) else True
        parity = {"count_matches_analytic": ok_count, "sampled_pairs_join_equal_keys": ok_keys}

    # ---- end to end through the host-buffer C-ABI call (rank-local; H2D + D2H inside the timed region) --------
    e2e = None
    if not args.no_e2e and args.workload != "c5":
        hR = torch.empty(b.n, dtype=b.dtype, pin_memory=True); hS = torch.empty(p.n, dtype=p.dtype, pin_memory=True)
        if world > 1:
            dist.broadcast(dR, src=0)
        hR.copy_(dR); hS.copy_(dS)
        cap = n_out[0]
        hOr = torch.empty(cap, dtype=torch.int32, pin_memory=True); hOs = torch.empty(cap, dtype=torch.int32, pin_memory=True)
        del last
        torch.cuda.empty_cache()
        times = []
        for i in range(1 + args.e2e_steps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t1 = time.perf_counter()
            got = lib.hjJoinHost(hR.data_ptr(), b.n, hS.data_ptr(), p.n, kb, hOr.data_ptr(), hOs.data_ptr(), cap)
            dt = time.perf_counter() - t1
            if got != cap:
                raise SystemExit(f"hjJoinHost returned {got}, expected {cap}: {lib.hjLastErrorString()}")
            if i > 0:
                times.append(dt)
        te = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": (nR_job + nS_job) / te.item(), "unit": UNIT, "h2d_bytes_per_step": (b.n + p.n) * kb, "d2h_bytes_per_step": cap * 8,
               "ms_per_step": te.item() * 1e3, "api": "hjJoinHost (include/hashjoin_b200.h), pinned host buffers"}
        lib.hashJoinRelease()

    # ---- the same step with the bucketised hash layout forced (C2's keys are a dense range; this is the generic path) ---
    hash_arm = None
    if args.layout == "auto" and not args.no_hash_arm and not (args.workload == "c5" and world > 1):
        lib.hjSetAllowDense(0)
        h_ms, _, h_ph, _, _ = run_timed(max(3, args.steps // 3), 3, False)
        hash_arm = {"value": (nR_job + nS_job) / (h_ms / 1e3), "unit": UNIT, "ms_per_step": h_ms,
                    "phases_ms": {k: sum(v) / len(v) for k, v in h_ph.items() if v},
                    "note": "hjSetAllowDense(0): same inputs, direct-address layout disabled"}
        if table_layout & 0x100:                              # the default counted by range test: show the match-cache variant of the same layout too
            lib.hjSetAllowDense(1)
            c_ms, _, c_ph, _, _ = run_timed(max(3, args.steps // 3), 3, False)
            hash_arm["direct_address_with_match_cache"] = {"value": (nR_job + nS_job) / (c_ms / 1e3), "unit": UNIT, "ms_per_step": c_ms,
                                                           "phases_ms": {k: sum(v) / len(v) for k, v in c_ph.items() if v},
                                                           "note": "hjSetAllowDense(1): direct-address table, lookups in the count pass, match cache"}
        lib.hjSetAllowDense(DENSE_POLICY[args.layout])

    # ---- single-pass probe (hjJoinFused): same build, then lookup + look-back + write in one kernel into a result of |S| pairs ---
    fused_arm = None
    if args.layout == "auto" and not args.no_hash_arm and not (args.workload == "c5" and world > 1) and cfg.expected_out is not None and cfg.expected_out <= p.n:
        fR, fS = result_columns(p.n)
        f_ms = []
        for i in range(3 + max(3, args.steps // 3)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if world > 1:
                dist.broadcast(dR, src=0)
            join.buildTable(dR, table)
            nf = join.join_fused(dS, table, fR, fS, probeRowBase=plo)
            e1.record(stream)
            torch.cuda.synchronize()
            if i >= 3:
                f_ms.append(e0.elapsed_time(e1))
        if nf != n_out[0]:
            raise SystemExit(f"hjJoinFused found {nf} pairs, count + write found {n_out[0]}")
        tf = torch.tensor([sum(f_ms) / len(f_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        fused_arm = {"value": (nR_job + nS_job) / (tf.item() / 1e3), "unit": UNIT, "ms_per_step": tf.item(),
                     "note": "hjJoinFused: build + ONE probe pass (lookup, decoupled look-back, write) into a caller-bounded result; "
                             "not available to the reference's count -> allocate -> probe call sequence"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    out_per_gpu = n_out[0]
    ab = algorithmic_bytes(b.n, p.n if args.workload != "c5" else p.n // world, out_per_gpu, kb,
                           table_in_hbm=lib.hjTableBytes(b.n, kb) > 96 * 2**20, lookups_in_write=bool(table_layout & 0x100))
    roofline = None
    c5_phases = None
    if "partition_exchange" in phase_ms and phase_ms["partition_exchange"]:
        # multi-GPU radix plan: the exchange is bound by NVLink, the rest by HBM. Rank 0's phases (the step time above is the max over ranks).
        px = sum(phase_ms["partition_exchange"]) / len(phase_ms["partition_exchange"]); lj = sum(phase_ms["local_join"]) / len(phase_ms["local_join"])
        sent = (world - 1) / world * (dR.numel() + dS.numel()) * (kb + 4)            # (key, global row id) tuples leaving this GPU
        c5_phases = {"partition_exchange_ms": px, "local_join_ms": lj, "nvlink_bytes_sent_per_gpu": int(sent),
                     "nvlink_achieved_gbs": sent / (px / 1e3) / 1e9, "nvlink_peak_gbs": 770.0,
                     "nvlink_frac": sent / (px / 1e3) / 1e9 / 770.0,
                     "note": "peak = measured peer copy per direction per GPU (B200_PROFILING.md); the phase also holds two histogram passes and the count-matrix all-gather"}
    if phase_ms.get("count"):
        k_ms = {k: sum(v) / len(v) for k, v in phase_ms.items()}
        dom = max(k_ms, key=k_ms.get)
        kernel = {"build": "build sequence (k_minmax, k_clear, k_build_dense | k_build_hash, ...)", "count": "k_count | k_count_sparse | k_count_range (+ k_sample_hits, k_scan_blocks and the 8-byte result-size readback)", "write": "k_write | k_write_sparse | k_write_range"}[dom]
        ach = ab[dom] / (k_ms[dom] / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "algorithmic_bytes": ab[dom], "kernel_ms": k_ms[dom], "peak_source": peak_src,
                    "phases_ms": k_ms, "share_of_step": k_ms[dom] / (sum(k_ms.values()) or 1),
                    "job": {"algorithmic_bytes": ab["total"], "achieved": ab["total"] / (ms_per_step / 1e3) / 1e9,
                            "frac": ab["total"] / (ms_per_step / 1e3) / 1e9 / peak}}
        traffic_file = ROOT / "profiles" / "traffic.json"
        if traffic_file.exists():
            roofline["traffic"] = json.loads(traffic_file.read_text()).get(dom if args.workload == "c2" and args.layout == "auto" else None)

    cpu = None
    if not args.no_cpu_baseline and world == 1:              # a reported baseline, timed on rank 0 at N = 1 only
        cpu = cpu_baseline(cfg, args.cpu_sample_log2)
        cpu.pop("seconds", None)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if (args.workload == "c5" and args.c5_total_log2) else "weak",
            "vs_baseline": None, "dtype": "int32" if kb == 4 else "int64", "data": "synthetic (seeded device generators, bit-identical to the oracle's)",
            "config": {"workload": workload_name(args, cfg), "build_rows": nR_job, "probe_rows": nS_job, "result_pairs": int(tot_out.item()),
                       "l2_hygiene": "inputs larger than L2 (probe column >= 1 GiB per GPU streams through every step)",
                       "timing": "CUDA events on the launching stream per step, summed over steps, max over ranks",
                       "wall_ms_per_step": wall_ms_per_step, "table_layout": args.layout,
                       "table_layout_chosen": {0: "bucketised hash", 1: "direct-address", 2: "grouped"}.get(table_layout & 0xFF, "?") + (", count by range test" if table_layout & 0x100 else "") + (", hit lists" if hit_lists else "")},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": timed_launches, "clocks": clocks, "parity": parity, "hash_layout": hash_arm, "fused_single_pass": fused_arm}
    if c5_phases is not None:
        line["c5_phases"] = c5_phases
        line["parity"] = {"count_matches_analytic": int(tot_out.item()) == cfg.expected_out} if cfg.expected_out is not None else None
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
