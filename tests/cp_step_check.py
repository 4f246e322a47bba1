"""Context-parallel inference_single_step vs the CPU oracle, repeated, with and without the audio side stream.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/cp_step_check.py [repeats]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

import mova_oracle as O
from cp_check import Mesh1D
from util import bf16_round, build_step_towers, metrics


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    from dualforce_b200 import pipeline as pl
    from dualforce_b200 import step

    cfg = dict(O.TINY_STEP_CFG, visual_dim=512, visual_heads=4, visual_ffn=768, grid_size=(3, 3, 5))
    if cfg["visual_heads"] % world:
        cfg = dict(cfg, visual_heads=world, visual_dim=128 * world, visual_ffn=256 * world)
    Pv, Pa, Pb, inp = O.make_step_case(cfg, 77)
    Pv, Pa, Pb = bf16_round(Pv), bf16_round(Pa), bf16_round(Pb)
    ctx = inp["context"].to(torch.bfloat16)
    rv, ra = O.inference_single_step(Pv, Pa, Pb, cfg, inp["visual_latents"], inp["audio_latents"], ctx.float(),
                                     inp["timestep"])
    try:
        for side in (True, False, True):
            pl.CPRuntime.audio_side_stream = side
            pl._RUNTIMES.clear()
            for rep in range(repeats):
                vis, aud, bridge, pipe = build_step_towers(cfg, Pv, Pa, Pb)  # cold memos every time
                kw = dict(visual_dit=vis, visual_latents=inp["visual_latents"].cuda(),
                          audio_latents=inp["audio_latents"].cuda(), context=ctx.cuda(), timestep=inp["timestep"].cuda(),
                          audio_timestep=None, video_fps=cfg["video_fps"])
                v, a = pipe.inference_single_step(**kw, cp_mesh=Mesh1D(dist.group.WORLD, rank, world))
                v2, a2 = pipe.inference_single_step(**kw, cp_mesh=Mesh1D(dist.group.WORLD, rank, world))  # warm memos
                v1, a1 = pipe.inference_single_step(**kw)  # cp = 1 on this rank
                torch.cuda.synchronize()
                mv, ma, ma2, ma1 = metrics(v, rv), metrics(a, ra), metrics(a2, ra), metrics(a1, ra)
                print(f"rank {rank} side_stream={side} rep {rep}: visual fro {mv['rel_fro']:.4f} | audio fro cold "
                      f"{ma['rel_fro']:.4f} warm {ma2['rel_fro']:.4f} cp1 {ma1['rel_fro']:.4f}", flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
