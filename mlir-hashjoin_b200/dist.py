"""Multi-GPU plans (one process per GPU, torch.distributed over NCCL/NVLink). The reference has none of this
(projectDescription.md:24 lists partitioned joins as left out); SURVEY.md section 8e defines the two plans.

broadcast build   probe rows range-sharded, build relation replicated: one broadcast of the build columns, then every
                  rank runs the single-GPU join on its shard with probe_row = shard base + local row. No exchange of S.
radix partition   both relations partitioned on an independent key hash by K5 (hjPartition), count matrix all-gathered,
                  (key, global row id) exchanged with a variable-count all-to-all, local join with payload columns.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _lib, join


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row range [lo, hi) of rank ``rank``: sizes differ by at most one row."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_build_join(build_keys: torch.Tensor, probe_shard: torch.Tensor, probe_row_base: int, src: int = 0,
                         group=None, table: join.HashTable | None = None, replicated: bool = False):
    """Every rank ends with the pairs of ITS probe rows; the global result is the concatenation (a multiset).
    ``build_keys`` must be allocated at full size on every rank; its contents matter on ``src`` only unless
    ``replicated`` says every rank already holds it."""
    if not replicated and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(build_keys, src=src, group=group)
    return join.hash_join(build_keys, probe_shard, table=table, probeRowBase=probe_row_base)


def partition(keys: torch.Tensor, row_base: int, n_parts: int, rows: torch.Tensor | None = None):
    """K5 on this rank's shard. Returns (keys_by_part, rows_by_part, offsets[n_parts+1] on the device)."""
    lib = _lib.load()
    kb = keys.element_size()
    n = keys.numel()
    out_keys = torch.empty_like(keys)
    out_rows = torch.empty(n, dtype=torch.int32, device=keys.device)
    offsets = torch.empty(n_parts + 1, dtype=torch.int64, device=keys.device)
    ws = torch.empty(lib.hjPartitionWorkspaceBytes(n, n_parts), dtype=torch.uint8, device=keys.device)
    rc = lib.hjPartition(keys.data_ptr(), None if rows is None else rows.data_ptr(), row_base & 0xFFFFFFFF, n, kb, n_parts,
                         out_keys.data_ptr(), out_rows.data_ptr(), offsets.data_ptr(), ws.data_ptr(), ws.numel(),
                         torch.cuda.current_stream().cuda_stream)
    _lib.check_status(rc, "hjPartition")
    return out_keys, out_rows, offsets


@dataclass
class ExchangePlan:
    send_counts: list[int]       # rows this rank sends to each peer
    recv_counts: list[int]       # rows this rank receives from each peer


def exchange_plan(offsets: torch.Tensor, group=None) -> ExchangePlan:
    """All-gather of the N x N count matrix (host-side plan for the variable-count all-to-all)."""
    world = dist.get_world_size(group)
    counts = (offsets[1:] - offsets[:-1]).to(torch.int64)
    gathered = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(gathered, counts, group=group)
    matrix = torch.stack(gathered).cpu()                       # matrix[src][dst]
    rank = dist.get_rank(group)
    return ExchangePlan(matrix[rank].tolist(), matrix[:, rank].tolist())


def exchange(parts: torch.Tensor, plan: ExchangePlan, group=None) -> torch.Tensor:
    """Variable-count all-to-all of a partition-ordered column. Works on any backend (NCCL on GPUs, gloo in CPU tests)."""
    out = torch.empty(sum(plan.recv_counts), dtype=parts.dtype, device=parts.device)
    dist.all_to_all_single(out, parts, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts, group=group)
    return out


@dataclass
class LandingPlan:
    """Where this rank's parts land in the owners' receive buffers (peer-store and copy-engine exchanges)."""
    received: int                # tuples this rank will hold once every rank's parts have landed
    fullest: int                 # tuples the fullest owner will hold (capacity check, identical on every rank)
    first_at_owner: list[int]    # per owner: first element of this rank's region in that owner's buffer (the lower ranks' parts precede it)
    offsets: list[int]           # per owner (+ end): this rank's parts in its own staging copy, grouped by owner


def landing_plan(matrix: torch.Tensor, rank: int) -> LandingPlan:
    """``matrix[src][dst]`` (host, int64) = tuples rank ``src`` sends to owner ``dst``. Senders fill an owner's buffer in rank order, so the
    regions tile [0, received) without gaps; pure host arithmetic, the same on every backend."""
    m = matrix.to(torch.int64)
    per_owner = m.sum(0)
    offsets = [0]
    for c in m[rank].tolist():
        offsets.append(offsets[-1] + int(c))
    return LandingPlan(int(per_owner[rank]), int(per_owner.max()) if per_owner.numel() else 0, [int(x) for x in m[:rank].sum(0).tolist()], offsets)


class PeerExchange:
    """Peer-mapped receive buffers for the fused partition + exchange (K5 ``hjPartitionPush``): every rank allocates the
    same symmetric buffers (torch symmetric memory = CUDA VMM handles exchanged once at rendezvous; plumbing only) and
    the partition kernel of each rank stores its tuples straight into the owner's buffer over NVLink. The only
    collective left on the data path is the all-gather of the N x N count matrix (N^2 int64)."""

    def __init__(self, capacity_rows: int, key_dtype: torch.dtype, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.group = group
        self.capacity = capacity_rows
        self.keys = symm_mem.empty(capacity_rows, dtype=key_dtype, device=device)
        self.rows = symm_mem.empty(capacity_rows, dtype=torch.int32, device=device)
        self.hk = symm_mem.rendezvous(self.keys, group)
        self.hr = symm_mem.rendezvous(self.rows, group)
        self.key_ptrs = torch.tensor(list(self.hk.buffer_ptrs), dtype=torch.int64, device=device)
        self.row_ptrs = torch.tensor(list(self.hr.buffer_ptrs), dtype=torch.int64, device=device)

    def barrier(self) -> None:
        """Stream-ordered barrier across the ranks (signal pads in peer memory)."""
        self.hk.barrier()

    def count(self, keys: torch.Tensor):
        """Pass 1 of the fused exchange: tuples of ``keys`` per owner rank (device, int64[world]) and the workspace that keeps the
        per-block count matrix for push()."""
        lib = _lib.load()
        world = dist.get_world_size(self.group)
        kb, n = keys.element_size(), keys.numel()
        counts = torch.empty(world, dtype=torch.int64, device=keys.device)
        ws = torch.empty(lib.hjPartitionWorkspaceBytes(n, world), dtype=torch.uint8, device=keys.device)
        _lib.check_status(lib.hjPartitionCount(keys.data_ptr(), n, kb, world, counts.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "hjPartitionCount")
        return counts, ws

    def push(self, keys: torch.Tensor, row_base: int, cursors: torch.Tensor, ws: torch.Tensor) -> None:
        """Pass 2: store every (key, row_base + i) into its owner's buffer, starting at ``cursors[owner]`` (device, int64[world]: the
        tuples the lower ranks send to that owner). Follows count() on the same keys and workspace; call barrier() afterwards."""
        lib = _lib.load()
        world = dist.get_world_size(self.group)
        rc = lib.hjPartitionPush(keys.data_ptr(), None, row_base & 0xFFFFFFFF, keys.numel(), keys.element_size(), world, self.key_ptrs.data_ptr(), self.row_ptrs.data_ptr(),
                                 cursors.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        _lib.check_status(rc, "hjPartitionPush")


    # ---- staged plan: partition into a local staging copy, copy engines carry the parts to their owners ----
    def scatter_local(self, keys: torch.Tensor, row_base: int, counts_row, ws: torch.Tensor, own_at: int) -> list[int]:
        """The push kernel with every destination on this GPU: (key, row_base + i) tuples of ``keys`` grouped by owner — part d at
        [offsets[d], offsets[d + 1]) of the staging buffers, except this rank's own part, which goes straight to its place in the
        receive buffer (from element ``own_at``). ``counts_row``: this rank's row of the count matrix (host ints). Follows count() on the
        same keys and workspace. Returns the offsets (host)."""
        lib = _lib.load()
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        n = keys.numel()
        if getattr(self, "stage_keys", None) is None or self.stage_keys.numel() < n:
            self.stage_keys = torch.empty(n, dtype=keys.dtype, device=keys.device)
            self.stage_rows = torch.empty(n, dtype=torch.int32, device=keys.device)
            kp, rp = [self.stage_keys.data_ptr()] * world, [self.stage_rows.data_ptr()] * world
            kp[rank], rp[rank] = self.keys.data_ptr(), self.rows.data_ptr()
            self.stage_key_ptrs = torch.tensor(kp, dtype=torch.int64, device=keys.device)
            self.stage_row_ptrs = torch.tensor(rp, dtype=torch.int64, device=keys.device)
        offsets = [0]
        for c in counts_row:
            offsets.append(offsets[-1] + int(c))
        first = offsets[:-1]
        first[rank] = own_at
        cursors = torch.tensor(first, dtype=torch.int64).to(keys.device, non_blocking=True)
        rc = lib.hjPartitionPush(keys.data_ptr(), None, row_base & 0xFFFFFFFF, n, keys.element_size(), world, self.stage_key_ptrs.data_ptr(), self.stage_row_ptrs.data_ptr(),
                                 cursors.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        _lib.check_status(rc, "hjPartitionPush")
        return offsets

    def send(self, offsets: list[int], first_at_owner: list[int]) -> None:
        """Device-to-device copies (copy engines, current stream) of every staged part into its owner's receive buffer at
        ``first_at_owner[d]``. Owners are visited starting with the next rank, so that at any time every owner receives from one sender."""
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if getattr(self, "peer_views", None) is None:
            self.peer_views = [(self.hk.get_buffer(d, (self.capacity,), self.keys.dtype), self.hr.get_buffer(d, (self.capacity,), torch.int32)) for d in range(world)]
        for step in range(1, world):                                # (this rank's own part is already in place: scatter_local)
            d = (rank + step) % world
            a, z = offsets[d], offsets[d + 1]
            if z > a:
                pk, pr = self.peer_views[d]
                at = first_at_owner[d]
                pk[at:at + (z - a)].copy_(self.stage_keys[a:z], non_blocking=True)
                pr[at:at + (z - a)].copy_(self.stage_rows[a:z], non_blocking=True)


def radix_join_staged(build_shard: torch.Tensor, build_row_base: int, probe_shard: torch.Tensor, probe_row_base: int,
                      build_x: PeerExchange, probe_x: PeerExchange, exchanged=None, table: join.HashTable | None = None, marks: dict | None = None,
                      result=None):
    """Radix-partitioned join whose exchange runs on the COPY ENGINES: each relation is partitioned by owner into a local staging copy
    (the push kernel, every destination local), the parts cross NVLink as device-to-device copies on a second stream — 770 GB/s per
    direction beside SM work (tools/peer_bench.py) — and the SMs go on meanwhile: the probe side is partitioned while the build side
    travels, the received build side is partitioned for the local join (hjBuild) while the probe side travels. One collective (count
    rows) and one host read per relation; a receive buffer that would overflow raises HashJoinError on every rank alike (radix_join()
    is the fallback; when it is the probe side that overflows, the build side has been exchanged already — into buffers nobody reads). ``marks``: CUDA events after each leg, both streams. ``result``:
    optional callable n -> (outR, outS) int32 tensors of n elements (a caller that keeps its result columns across joins)."""
    group = build_x.group
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if table is None:
        table = join.allocateHashTable(build_x.capacity, None, build_shard.dtype, build_shard.device)
    main = torch.cuda.current_stream()
    ce = _side_stream(build_shard.device)

    def mark(name, stream=None):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream if stream is not None else main)
            marks[name] = ev
    build_x.barrier()                                               # nobody is still reading last step's buffers
    mark("start")
    staged_b, landed_b, staged_p, landed_p = (torch.cuda.Event() for _ in range(4))

    def plan_of(x: PeerExchange, shard: torch.Tensor, side: str):
        """Histogram, all-gather of the count rows, host read (one sync per relation: the build side's parts are on the links while the
        probe side is still being counted)."""
        counts, ws = x.count(shard)
        mark("histogram_" + side)
        matrix = torch.empty(world, world, dtype=torch.int64, device=shard.device)
        dist.all_gather_into_tensor(matrix.view(-1), counts, group=group)
        m = matrix.cpu()
        land = landing_plan(m, rank)
        if land.fullest > x.capacity:                               # the same verdict on every rank
            raise _lib.HashJoinError("receive buffer too small for this key distribution (skew): use radix_join()")
        mark("count_matrix_" + side)
        return m, ws, land

    mb, wsb, lb = plan_of(build_x, build_shard, "build")
    off_b = build_x.scatter_local(build_shard, build_row_base, mb[rank].tolist(), wsb, lb.first_at_owner[rank])
    staged_b.record(main)
    mark("scatter_build")
    with torch.cuda.stream(ce):
        ce.wait_event(staged_b)
        mark("copy_build_start", ce)
        build_x.send(off_b, lb.first_at_owner)
        build_x.barrier()                                           # every rank's build tuples have landed
        landed_b.record(ce)
        mark("copy_build_end", ce)
    mp, wsp, lp = plan_of(probe_x, probe_shard, "probe")
    off_p = probe_x.scatter_local(probe_shard, probe_row_base, mp[rank].tolist(), wsp, lp.first_at_owner[rank])
    staged_p.record(main)
    mark("scatter_probe")
    with torch.cuda.stream(ce):
        ce.wait_event(staged_p)
        mark("copy_probe_start", ce)
        probe_x.send(off_p, lp.first_at_owner)
        probe_x.barrier()                                           # every rank's probe tuples have landed
        landed_p.record(ce)
        mark("copy_probe_end", ce)
    nb, npr = lb.received, lp.received
    main.wait_event(landed_b)
    join.initializeHashTable(table)
    join.buildTable(build_x.keys[:nb], table, build_x.rows[:nb])    # beside the probe side's copies
    mark("local_build")
    main.wait_event(landed_p)
    if exchanged is not None:
        exchanged()
    pk, pr = probe_x.keys[:npr], probe_x.rows[:npr]
    n = join.countRows(pk, table, pr, 0)
    mark("local_count")
    if result is not None:
        outR, outS = result(n)
    else:
        outR = torch.empty(n, dtype=torch.int32, device=pk.device)
        outS = torch.empty(n, dtype=torch.int32, device=pk.device)
    if n:
        join.probeRelation(pk, table, outR, outS, pr, 0)
    mark("local_write")
    return outR, outS


def exchange_fused(build_shard: torch.Tensor, build_row_base: int, probe_shard: torch.Tensor, probe_row_base: int,
                   build_x: PeerExchange, probe_x: PeerExchange, marks: dict | None = None, build_landed: torch.cuda.Event | None = None) -> tuple[int, int]:
    """Both relations through the fused partition + exchange with ONE collective (the all-gather of both count rows) and ONE host
    readback (the count matrices, needed to size the local join). Returns the tuples this rank receives (build, probe); they are all
    there after the closing barrier. Raises HashJoinError — on every rank alike — when an owner would receive more than its buffer
    holds (skew beyond the slack): nothing has been stored then, and radix_join() (NCCL all-to-all, exact sizes) is the fallback.
    ``marks``: filled with CUDA events after each leg (bench.py times the legs). ``build_landed``: recorded once the build side has
    landed on every rank, before the probe side is pushed (another stream can wait on it and start the local build)."""
    group = build_x.group
    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def mark(name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream())
            marks[name] = ev
    build_x.barrier()                                               # nobody is still reading last step's buffers
    mark("start")
    cb, wsb = build_x.count(build_shard)
    cp, wsp = probe_x.count(probe_shard)
    mark("histograms")
    matrix = torch.empty(world, 2 * world, dtype=torch.int64, device=build_shard.device)
    dist.all_gather_into_tensor(matrix.view(-1), torch.cat([cb, cp]), group=group)      # matrix[src] = [build counts per dst | probe counts per dst]
    host = matrix.cpu()                                             # the step's one host sync
    mb, mp = host[:, :world], host[:, world:]
    lb, lp = landing_plan(mb, rank), landing_plan(mp, rank)
    if lb.fullest > build_x.capacity or lp.fullest > probe_x.capacity:
        raise _lib.HashJoinError("receive buffer too small for this key distribution (skew): use radix_join()")
    cursors = torch.tensor(lb.first_at_owner + lp.first_at_owner, dtype=torch.int64).to(build_shard.device, non_blocking=True)   # first element of my region in every owner's buffer
    mark("count_matrix")
    nb, npr = lb.received, lp.received
    build_x.push(build_shard, build_row_base, cursors[:world], wsb)
    build_x.barrier()                                               # every rank's build tuples have landed
    mark("push_build")
    if build_landed is not None:
        build_landed.record(torch.cuda.current_stream())
    probe_x.push(probe_shard, probe_row_base, cursors[world:], wsp)
    probe_x.barrier()                                               # every rank's probe tuples have landed
    mark("push_probe")
    return nb, npr


def radix_join_fused(build_shard: torch.Tensor, build_row_base: int, probe_shard: torch.Tensor, probe_row_base: int,
                     build_x: PeerExchange, probe_x: PeerExchange, exchanged=None, table: join.HashTable | None = None,
                     marks: dict | None = None, overlap_build: bool = False, result=None):
    """Radix-partitioned join with the exchange fused into the partition kernel (peer stores over NVLink).
    ``exchanged`` (optional callable) runs once every rank's tuples have landed, before the probe passes of the local join (bench.py
    marks a phase there). ``overlap_build``: the local build (hjBuild on the received build tuples) runs on a second stream while the
    probe side is still crossing NVLink — the push is bound by the links, the build by HBM."""
    if table is None:
        table = join.allocateHashTable(build_x.capacity, None, build_shard.dtype, build_shard.device)
    main = torch.cuda.current_stream()
    landed = torch.cuda.Event() if overlap_build else None
    nb, npr = exchange_fused(build_shard, build_row_base, probe_shard, probe_row_base, build_x, probe_x, marks, build_landed=landed)
    if overlap_build:                                               # everything above is queued, nothing has been waited for: the probe push runs on `main` ...
        side = _side_stream(build_shard.device)
        with torch.cuda.stream(side):                               # ... while the build of the received build tuples runs here
            side.wait_event(landed)
            join.buildTable(build_x.keys[:nb], table, build_x.rows[:nb])
        main.wait_stream(side)
    if exchanged is not None:
        exchanged()
    def mark(name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(main)
            marks[name] = ev
    if not overlap_build:
        join.initializeHashTable(table)
        join.buildTable(build_x.keys[:nb], table, build_x.rows[:nb])
    mark("local_build")
    pk, pr = probe_x.keys[:npr], probe_x.rows[:npr]
    n = join.countRows(pk, table, pr, 0)
    mark("local_count")
    if result is not None:
        outR, outS = result(n)
    else:
        outR = torch.empty(n, dtype=torch.int32, device=pk.device)
        outS = torch.empty(n, dtype=torch.int32, device=pk.device)
    if n:
        join.probeRelation(pk, table, outR, outS, pr, 0)
    mark("local_write")
    return outR, outS


_SIDE_STREAMS: dict = {}


def _side_stream(device) -> torch.cuda.Stream:
    key = str(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def radix_join(build_shard: torch.Tensor, build_row_base: int, probe_shard: torch.Tensor, probe_row_base: int, group=None, table: join.HashTable | None = None):
    """Radix-partitioned join of range-sharded relations. Returns this rank's (build_row, probe_row) pairs, global row ids."""
    world = dist.get_world_size(group)
    bk, br, bo = partition(build_shard, build_row_base, world)
    pk, pr, po = partition(probe_shard, probe_row_base, world)
    bplan, pplan = exchange_plan(bo, group), exchange_plan(po, group)
    my_bk, my_br = exchange(bk, bplan, group), exchange(br, bplan, group)
    my_pk, my_pr = exchange(pk, pplan, group), exchange(pr, pplan, group)
    return join.hash_join(my_bk, my_pk, table=table, buildPayload=my_br, probePayload=my_pr)
