"""NVLink push bandwidth by store shape (tools/peer_bench.cu): every rank streams 1 GiB from its HBM into the next rank's symmetric buffer.
Run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_bench.py"""
import ctypes
import json
import os
from pathlib import Path

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = ctypes.CDLL(str(Path(__file__).resolve().parent / "libpeer_bench.so"))
    lib.peer_push.restype = ctypes.c_int
    lib.peer_push.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    n = 1 << 30
    buf = symm_mem.empty(n, dtype=torch.uint8, device=dev)
    h = symm_mem.rendezvous(buf, dist.group.WORLD)
    src = torch.randint(0, 255, (n,), dtype=torch.uint8, device=dev)
    peer = h.buffer_ptrs[(rank + 1) % world]
    res = {}
    for target, tname in ((peer, "peer"), (buf.data_ptr(), "local")):
        for mode, name in ((0, "4B/lane"), (1, "8B/lane"), (2, "16B/lane"), (3, "TMA bulk 4KB")):
            for ctas in (148 * 2, 148 * 4, 148 * 8):
                best = 1e9
                for rep in range(4):
                    torch.cuda.synchronize(); dist.barrier()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    rc = lib.peer_push(src.data_ptr(), target, n, mode, ctas, torch.cuda.current_stream().cuda_stream)
                    assert rc == 0, rc
                    b.record(); b.synchronize()
                    best = min(best, a.elapsed_time(b))
                res[f"{tname} {name} x{ctas}"] = round(n / best / 1e6, 1)
        h.barrier()
        torch.cuda.synchronize()
    # copy engine: the same 1 GiB as one device-to-device copy into the peer's buffer — alone, and while the SMs are kept busy by a local
    # streaming kernel on another stream (does the copy run beside SM work?)
    peer_t = h.get_buffer((rank + 1) % world, (n,), torch.uint8)
    local_dst = torch.empty(n, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream()
    for busy in (False, True):
        best, best_k = 1e9, 1e9
        for rep in range(4):
            torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if busy:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    ka.record(side)
                    for _ in range(8):
                        lib.peer_push(src.data_ptr(), local_dst.data_ptr(), n, 2, 148 * 8, side.cuda_stream)
                    kb.record(side)
            a.record()
            peer_t.copy_(src, non_blocking=True)
            b.record(); b.synchronize()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
            if busy:
                best_k = min(best_k, ka.elapsed_time(kb) / 8)
        res["peer copy_ (1 GiB)" + (" beside a local streaming kernel" if busy else "")] = round(n / best / 1e6, 1)
        if busy:
            res["local 16B/lane kernel beside the peer copy"] = round(n / best_k / 1e6, 1)
    h.barrier()
    torch.cuda.synchronize()
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        print(json.dumps({"ranks": world, "gbs_rank0": out[0], "gbs_min_over_ranks": {k: min(r[k] for r in out) for k in out[0]}}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
