#!/usr/bin/env python
"""Do copies overlap a kernel that owns every SM?  (two ranks, one per GPU)

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/ce_overlap_probe.py

The peer-memory Ulysses exchange (dualforce_b200/peer.py) relies on its transfers running on the copy engines while the
attention kernel holds all SMs.  This probe times, on a side stream, (a) a copy into the peer's IPC-mapped window,
(b) a same-device device-to-device copy, (c) an 8-byte peer copy, each alone and while a ~8 ms attention launch runs on
the main stream.  A copy that needs SMs finishes only when the attention kernel drains."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dualforce_b200 import ops, peer

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nbytes = 160 << 20
    win = peer.CudaIpcWindow(dist.group.WORLD, rank, world, dev)
    win.ensure(peer.FLAG_BYTES + 2 * nbytes)
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev).fill_(rank + 1)
    other = (rank + 1) % world
    S, H = 43120, 10
    q = torch.randn(1, S, H * 128, device=dev, dtype=torch.bfloat16)
    side = torch.cuda.Stream()
    lib = win._lib.load()
    import ctypes

    def copy(dst_ptr, n):
        def run():
            win._lib.check(lib.mova_b200_peer_push(1, (ctypes.c_void_p * 1)(dst_ptr), (ctypes.c_void_p * 1)(src.data_ptr()),
                                                   (ctypes.c_int64 * 1)(n), 0, (ctypes.c_void_p * 1)(), 0, 1, None,
                                                   side.cuda_stream), "push", launches=0)
        return run

    cases = {"peer_copy_160MB": copy(win.ptrs[other] + peer.FLAG_BYTES, nbytes),
             "self_copy_160MB": copy(win.ptrs[rank] + peer.FLAG_BYTES + nbytes, nbytes),
             "peer_copy_8B": copy(win.ptrs[other] + 4096, 8),
             "self_copy_8B": copy(win.ptrs[rank] + 4096 + 64, 8)}
    out = {}
    for name, fn in cases.items():
        for busy in (False, True):
            ts = []
            for it in range(4):
                torch.cuda.synchronize()
                dist.barrier()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                if busy:
                    ops.attention(q, q, q, H)
                a1.record()
                side.wait_event(a0)  # the copy is queued at the same time the attention kernel is
                with torch.cuda.stream(side):
                    c0.record()
                    fn()
                    c1.record()
                torch.cuda.synchronize()
                ts.append((round(c0.elapsed_time(c1), 4), round(a0.elapsed_time(c1), 4), round(a0.elapsed_time(a1), 4)))
            out[f"{name}{'+attention' if busy else ''}"] = {"copy_ms": ts[-1][0], "copy_done_after_ms": ts[-1][1],
                                                             "attention_ms": ts[-1][2], "all": ts}
    res = [None] * world
    dist.all_gather_object(res, out)
    if rank == 0:
        print(json.dumps({"probe": "copy engines vs a kernel that owns every SM", "ranks": res}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
