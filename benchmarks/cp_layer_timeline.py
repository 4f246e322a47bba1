#!/usr/bin/env python
"""Per-segment device timeline of the context-parallel forward (the stand-in for an nsys trace, which this image does not
have): CUDA events around every segment of every layer, on the stream the segment runs on.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/cp_layer_timeline.py [--single-stream]
        [--policies peer:1,3,1 peer:1,4 peer:1,1,1,1,1 nccl:1,1,1,1,1 ...]

`--policies` measures several (exchange data path : attention set sizes) configurations in ONE process, one JSON line
each (model construction dominates the run time of this tool); without it the defaults of the build are measured.
`forward_ms_plain` is the mean of 3 forwards without any timeline events.

Prints one JSON object per policy (rank 0): per segment name the mean milliseconds per layer and the sum over one forward, per
stream; `main` segments add up to the critical path of the rank, `comm` / `audio` segments run beside it.  The two
`wait_all_to_all_*` segments are the EXPOSED part of the Ulysses exchanges (time the main stream idles for them)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--single-stream", action="store_true", help="audio tower + v2a on the main stream (round-1 order)")
    ap.add_argument("--policies", nargs="*", default=["default"],
                    help="exchange:set_sizes, e.g. peer:1,3,1  nccl:1,1,1,1,1  (sets only matter for an odd head count)")
    args = ap.parse_args()
    import bench
    from dualforce_b200 import _lib, pipeline as pl
    from torch.distributed.device_mesh import init_device_mesh

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    mesh = init_device_mesh("cuda", (world,), mesh_dim_names=("cp",))
    pl.CPRuntime.audio_side_stream = not args.single_stream
    cfg = dict(bench.FULL_360P)
    pipe = bench.build_model(cfg, device, experts=1)
    host = bench.host_step_inputs(cfg, pin=False)
    x_in = torch.cat([host["latents"], host["condition"]], dim=1).to(device)
    aud = host["audio_latents"].to(device)
    ctx = host["context_pos"].to(device)
    ts = torch.tensor([900.0], device=device)
    kw = dict(visual_dit=pipe.video_dit, visual_latents=x_in, audio_latents=aud, context=ctx, timestep=ts,
              audio_timestep=None, video_fps=cfg["video_fps"], cp_mesh=mesh)
    for policy in args.policies:
        if policy != "default":
            exch, _, sizes = policy.partition(":")
            exch, _, nstreams = exch.partition("@")  # peer@4 = remote pushes dealt over 4 streams
            pl.CPRuntime.push_streams_n = int(nstreams) if nstreams else 1
            for rt in pl._RUNTIMES.values():
                rt.push_streams_n = pl.CPRuntime.push_streams_n
            kernel_flags = exch == "peerk"  # peer windows with the kernel-based flag store / wait instead of memops
            exch = "peer" if kernel_flags else exch
            pl.CPRuntime.exchange = exch
            for rt in pl._RUNTIMES.values():
                rt.exchange = exch
                if rt._px:
                    rt._px.window.memops = (not kernel_flags) and bool(_lib.load().mova_b200_peer_memops_supported())
            pl.CPRuntime.set_sizes = tuple(int(v) for v in sizes.split(",")) if sizes else None
        for _ in range(2):
            pipe.inference_single_step(**kw)
        torch.cuda.synchronize()
        dist.barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(3):
            pipe.inference_single_step(**kw)
        p1.record()
        torch.cuda.synchronize()
        plain = torch.tensor([p0.elapsed_time(p1) / 3], device=device)
        dist.all_reduce(plain, op=dist.ReduceOp.MAX)
        dist.barrier()
        pl.TIMELINE = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.inference_single_step(**kw)
        e1.record()
        torch.cuda.synchronize()
        tl, pl.TIMELINE = pl.TIMELINE, None
        agg = {}
        for name, stream, a, b in tl:
            key = f"{stream}:{name}"
            d = agg.setdefault(key, [0.0, 0])
            d[0] += a.elapsed_time(b)
            d[1] += 1
        if rank == 0:
            used = sorted({(("peer/memops" if rt._px.window.memops else "peer/kernel-flags")
                            if (rt._px and rt.exchange == "peer") else "nccl") for rt in pl._RUNTIMES.values()})
            out = {"cp": world, "policy": policy, "exchange_used": used, "single_stream": bool(args.single_stream),
                   "forward_ms_plain": float(plain.item()), "forward_ms": e0.elapsed_time(e1),
                   "segments": {k: {"total_ms": round(v[0], 3), "count": v[1], "mean_ms": round(v[0] / v[1], 4)}
                                for k, v in sorted(agg.items())},
                   "main_stream_total_ms": round(sum(v[0] for k, v in agg.items() if k.startswith("main:")), 3),
                   "note": "events add a little launch overhead; forward_ms is therefore slightly above the bench's"}
            print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
