"""2-GPU microbenchmark of the halo transports (run under torchrun, 2 ranks): bytes moved per direction per
second for (a) a copy-engine peer copy, (b) pc_halo_push on a contiguous run, (c) pc_halo_push through a random
index, (d) NCCL all_to_all_single; each with both ranks sending at once and with rank 0 sending alone."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm
    from pcompanion_b200._lib import call, dev as dptr, stream
    rows, width = 1_000_000, 256
    nbytes = rows * width * 4
    buf = symm.empty(rows * width, dtype=torch.float32, device=dev)
    h = symm.rendezvous(buf, dist.group.WORLD.group_name)
    peer = 1 - rank
    peer_view = h.get_buffer(peer, (rows, width), torch.float32)
    src = torch.randn(rows, width, device=dev)
    idx = torch.randperm(rows, device=dev)
    recv = torch.empty_like(src)
    off = (ctypes.c_int64 * 3)(*([0, 0, rows] if rank == 0 else [0, rows, rows]))
    base = (ctypes.c_void_p * 2)(*[int(h.buffer_ptrs[p]) for p in range(2)])
    zero = (ctypes.c_int64 * 2)(0, 0)
    flag = torch.zeros(1, device=dev)

    def push(index):
        call("pc_halo_push", dptr(src, torch.float32, "src"), width, None if index is None else dptr(index, torch.int64, "i"),
             2, off, base, zero, zero, 0, width, stream())

    def a2a():
        dist.all_to_all_single(recv, src, output_split_sizes=[0, rows] if rank == 0 else [rows, 0],
                               input_split_sizes=[0, rows] if rank == 0 else [rows, 0])

    variants = {"copy_engine": lambda: peer_view.copy_(src), "push_contiguous": lambda: push(None),
                "push_indexed": lambda: push(idx), "nccl_all_to_all": a2a}
    out = {}
    for name, fn in variants.items():
        for mode in ("both", "rank0_only"):
            if name == "nccl_all_to_all" and mode == "rank0_only":
                continue
            active = mode == "both" or rank == 0
            for _ in range(3):
                if active:
                    fn()
            torch.cuda.synchronize(); dist.all_reduce(flag); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                if active:
                    fn()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            out[f"{name}/{mode}"] = {"ms": round(ms.item(), 3), "GBps_per_direction": round(nbytes / ms.item() / 1e6, 1)}
            dist.all_reduce(flag); torch.cuda.synchronize()
    ok = bool(torch.equal(h.get_buffer(rank, (rows, width), torch.float32)[:8], recv[:8])) if False else None
    if rank == 0:
        print(json.dumps({"rows": rows, "bytes": nbytes, "results": out}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
