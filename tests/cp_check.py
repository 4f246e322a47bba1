"""Context-parallel forward vs the CPU oracle.  Run under torchrun (one rank per GPU) or through
tests/test_gpu_cp.py:   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/cp_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

import mova_oracle as O
from util import assert_close, bf16_round, build_towers, to_dev


class Mesh1D:
    """The three methods forward_dual_tower_dit uses from a 1-D DeviceMesh slice (pipeline_mova.py:654-656)."""

    def __init__(self, group, rank, size):
        self._g, self._r, self._s = group, rank, size

    def get_group(self):
        return self._g

    def get_local_rank(self):
        return self._r

    def size(self):
        return self._s


CP_CFG = dict(O.TINY_CFG, visual_dim=512, visual_heads=4, visual_ffn=768, grid_size=(3, 3, 5), audio_len=21)


def run_check(rank: int, world: int, cfg=None, seed: int = 77):
    if cfg is None:
        cfg = CP_CFG
        if cfg["visual_heads"] % world:  # e.g. 8 ranks: 8 video heads (1024 channels), 60 video tokens over 8 ranks
            cfg = dict(cfg, visual_heads=world, visual_dim=128 * world, visual_ffn=256 * world)
    torch.cuda.set_device(rank % torch.cuda.device_count())
    Pv, Pa, Pb, inp = O.make_case(cfg, seed)
    Pv, Pa, Pb, inp = bf16_round(Pv), bf16_round(Pa), bf16_round(Pb), bf16_round(inp)
    vis, aud, bridge, pipe = build_towers(cfg, Pv, Pa, Pb)
    d = to_dev(inp)
    mesh = Mesh1D(dist.group.WORLD, rank, world)
    fv, fa = pipe.forward_dual_tower_dit(vis, d["visual_x"], d["audio_x"], d["visual_context"], d["audio_context"],
                                         d["visual_t_mod"], d["audio_t_mod"], d["visual_freqs"], d["audio_freqs"],
                                         cfg["grid_size"], cfg["video_fps"], cp_mesh=mesh)
    torch.cuda.synchronize()
    rv, ra = O.forward_dual_tower_dit(Pv, Pa, Pb, cfg, inp["visual_x"], inp["audio_x"], inp["visual_context"],
                                      inp["audio_context"], inp["visual_t_mod"], inp["audio_t_mod"], inp["visual_freqs"],
                                      inp["audio_freqs"], cfg["grid_size"], cfg["video_fps"])
    assert fv.shape == rv.shape and fa.shape == ra.shape  # full-length outputs on every rank
    mv = assert_close(fv, rv, f"cp{world} visual (rank {rank})", ratio=3e-2, fro=1.2e-2)
    ma = assert_close(fa, ra, f"cp{world} audio (rank {rank})", ratio=3e-2, fro=1.2e-2)
    # and against the same modules at cp = 1 on this GPU: only bf16 re-association noise may differ
    sv, sa = pipe.forward_dual_tower_dit(vis, d["visual_x"], d["audio_x"], d["visual_context"], d["audio_context"],
                                         d["visual_t_mod"], d["audio_t_mod"], d["visual_freqs"], d["audio_freqs"],
                                         cfg["grid_size"], cfg["video_fps"])
    assert_close(fv, sv.float().cpu(), "cp vs single-GPU visual", ratio=2e-2, fro=6e-3)
    assert_close(fa, sa.float().cpu(), "cp vs single-GPU audio", ratio=2e-2, fro=6e-3)
    # the attention-processor level drop-in (reference USPAttention contract: sequence shards in and out)
    import dualforce_b200 as B

    H = max(world, 2) if (max(world, 2) % world == 0) else world
    S = 64 * world
    g = torch.Generator().manual_seed(5)
    q, k, v = (torch.randn(1, S, H * 128, generator=g).to(torch.bfloat16) for _ in range(3))
    ref = O.attention(q.float(), k.float(), v.float(), H)
    sl = S // world
    sh = slice(rank * sl, (rank + 1) * sl)
    out = B.USPAttention(H)(q[:, sh].cuda().contiguous(), k[:, sh].cuda().contiguous(), v[:, sh].cuda().contiguous())
    assert_close(out, ref[:, sh], f"USPAttention world {world}", ratio=1e-2, fro=6e-3)
    return mv, ma


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))))
    from dualforce_b200 import pipeline as pl

    try:
        # both data paths of the Ulysses exchange: copy-engine pushes into peer windows + flag words (the default),
        # and NCCL all_to_all_single
        for exchange in (sys.argv[1:] or ["peer", "nccl"]):
            pl.CPRuntime.exchange = exchange
            pl._RUNTIMES.clear()
            mv, ma = run_check(rank, world)
            used = sorted({("peer" if (rt._px and rt.exchange == "peer") else "nccl") for rt in pl._RUNTIMES.values()})
            if rank == 0:
                print(f"cp_check OK world={world} exchange={exchange} used={used}", {"visual": mv, "audio": ma},
                      flush=True)
            # an odd number of heads per rank (like the 5 of MOVA-360p at cp = 8): exchanged head by head, attended in
            # sets of head groups -- 3 heads per rank here ([[0], [1], [2]] on two side streams)
            odd = dict(CP_CFG, visual_heads=3 * world, visual_dim=384 * world, visual_ffn=512 * world)
            mv, ma = run_check(rank, world, cfg=odd, seed=78)
            if rank == 0:
                print(f"cp_check OK world={world} exchange={exchange} (3 heads per rank)", {"visual": mv, "audio": ma},
                      flush=True)
            if exchange == "peer" and used != ["peer"]:
                raise SystemExit(f"rank {rank}: the peer-memory exchange was requested but {used} ran")
        pl.CPRuntime.exchange = "peer"
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
