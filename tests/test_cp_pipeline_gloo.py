"""The REAL context-parallel host path (dualforce_b200.pipeline / step with a cp_mesh) on CPU: world_size 2 over gloo,
kernels replaced by tests/emulated_ops.py.  Covers what test_cp_gloo.py (hand-written layout walk) cannot: the
product's own `_forward_eager` CP branch, weight permutation cache, sharded a2v / v2a bridge with LSE merge, the
replicated audio tower, and the step wrapper's sequence-sharded head + narrow all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mova_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Mesh:
    """The three methods of a 1-D DeviceMesh slice the path uses (pipeline_mova.py:654-656)."""

    def __init__(self, rank, world):
        self._rank, self._world = rank, world

    def get_group(self):
        return None  # default (world) group

    def get_local_rank(self):
        return self._rank

    def size(self):
        return self._world


def _worker(rank, world, port, grid, heads, results, exchange="collective", set_sizes=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dualforce_b200.ops as real_ops
        import emulated_ops
        from util import bf16_round, build_step_towers, metrics

        for name, fn in emulated_ops.ENTRY_POINTS.items():
            setattr(real_ops, name, fn)
        cfg = dict(O.TINY_STEP_CFG, grid_size=grid)
        if heads is not None:  # e.g. 4 heads on 2 ranks: two head groups per rank, exchanged one after the other
            cfg.update(visual_heads=heads, visual_dim=128 * heads, visual_ffn=192 * heads)
        Pv, Pa, Pb, inp = O.make_step_case(cfg, 77)
        Pv, Pa, Pb = bf16_round(Pv), bf16_round(Pa), bf16_round(Pb)
        ctx = inp["context"].to(torch.bfloat16)
        vis, aud, bridge, pipe = build_step_towers(cfg, Pv, Pa, Pb, device="cpu")
        mesh = _Mesh(rank, world)
        window = None
        if exchange == "peer":
            # the product's peer-memory exchange (dualforce_b200/peer.py: offsets, flag words, epochs) over a
            # shared-memory stand-in for the CUDA IPC windows
            from dualforce_b200 import peer, pipeline as pl
            from shm_peer import ShmWindow

            rt = pl.CPRuntime.from_mesh(mesh, torch.device("cpu"))
            window = ShmWindow(rank, world, str(port), jitter_ms=3.0 * (1 + rank))  # ranks drift apart on purpose
            rt._px = peer.PeerExchange(window, rank, world)
            pl.CPRuntime.set_sizes = set_sizes
        kw = dict(visual_dit=vis, visual_latents=inp["visual_latents"], audio_latents=inp["audio_latents"], context=ctx,
                  timestep=inp["timestep"], audio_timestep=None, video_fps=cfg["video_fps"])
        v1, a1 = pipe.inference_single_step(**kw)                 # cp = 1 on this rank
        emulated_ops.CALLS.clear()
        v2, a2 = pipe.inference_single_step(**kw, cp_mesh=mesh)   # cp = 2, head on the local chunk
        calls = dict(emulated_ops.CALLS)
        same_as_collective = None
        if window is not None:  # the same forward through dist.all_to_all_single: must agree bit for bit
            rt.exchange = "nccl"
            vb, ab = pipe.inference_single_step(**kw, cp_mesh=mesh)
            rt.exchange = "peer"
            same_as_collective = bool(torch.equal(vb, v2) and torch.equal(ab, a2))
        rv, ra = O.inference_single_step(Pv, Pa, Pb, cfg, inp["visual_latents"], inp["audio_latents"], ctx.float(),
                                         inp["timestep"])
        # the forward-level API must still return full-length hidden states under CP
        from dualforce_b200 import step

        tok, g = step.patchify(vis, inp["visual_latents"])
        tok_a, (f,) = step.patchify(aud, inp["audio_latents"])
        t, t_mod = step.embed_time(vis, inp["timestep"])
        ta, ta_mod = step.embed_time(aud, inp["timestep"])
        full_v, full_a = pipe.forward_dual_tower_dit(
            vis, tok, tok_a, step.embed_text(vis, ctx), step.embed_text(aud, ctx), t_mod, ta_mod,
            step.token_freqs(vis, g, "cpu"), step.token_freqs(aud, (f,), "cpu"), g, cfg["video_fps"], cp_mesh=mesh)
        # the reference-style pad / gather helpers (functional.py:55-112 twins) across the two ranks
        from dualforce_b200 import cp

        tab = torch.arange(1 * 7 * 4, dtype=torch.float32).reshape(1, 7, 4)  # 7 rows: 4 + (3 + 1 pad) / 2 + 2 + 2 + (1 + 1)
        chunk, chunk_len, pad_len, total = cp._sp_split_tensor(tab, sp_size=world, sp_rank=rank)
        sp_ok = bool(torch.equal(cp._sp_all_gather_avg(chunk, sp_group=None, pad_len=pad_len), tab)
                     and (chunk_len, pad_len, total) == ((4, 1, 7) if world == 2 else (2, 1, 7)))
        # weights reloaded in place after a context-parallel forward: the permuted CP copies must follow
        Qv, Qa, Qb, _ = O.make_step_case(cfg, 78)
        Qv, Qa, Qb = bf16_round(Qv), bf16_round(Qa), bf16_round(Qb)
        vis.load_state_dict({k: v.to(torch.bfloat16) for k, v in Qv.items()})
        aud.load_state_dict({k: v.to(torch.bfloat16) for k, v in Qa.items()})
        bridge.load_state_dict({k: v.to(torch.bfloat16) for k, v in Qb.items()})
        v3, a3 = pipe.inference_single_step(**kw, cp_mesh=mesh)
        vis_f, aud_f, bridge_f, pipe_f = build_step_towers(cfg, Qv, Qa, Qb, device="cpu")
        v4, a4 = pipe_f.inference_single_step(**dict(kw, visual_dit=vis_f), cp_mesh=mesh)
        reload_ok = bool(torch.equal(v3, v4) and torch.equal(a3, a4) and not torch.equal(v3, v2))
        results[rank] = dict(sp_ok=sp_ok, reload_ok=reload_ok,
            cp_vs_oracle_v=metrics(v2, rv), cp_vs_oracle_a=metrics(a2, ra), cp_vs_cp1_v=metrics(v2, v1.float()),
            cp_vs_cp1_a=metrics(a2, a1.float()), shapes=(tuple(v2.shape), tuple(a2.shape), tuple(full_v.shape)),
            calls=calls, v2=v2.float(), a2=a2.float(),
            peer=None if window is None else dict(same_as_collective=same_as_collective, pushes=window.pushes, waits=window.waits, epoch=rt._px.epoch,
                                                  generation=window.generation))
    finally:
        dist.destroy_process_group()


def _check(results, world, grid, heads):
    assert len(results) == world
    f, h, w = grid
    dim = 256 if heads is None else 128 * heads
    for rank in range(world):
        r = results[rank]
        assert r["shapes"] == ((1, 16, f, 2 * h, 2 * w), (1, 32, 21), (1, f * h * w, dim))
        for key in ("cp_vs_oracle_v", "cp_vs_oracle_a"):
            m = r[key]
            assert m["finite"] and m["ratio"] <= 3e-2 and m["rel_fro"] <= 1.5e-2, (rank, key, m)
        for key in ("cp_vs_cp1_v", "cp_vs_cp1_a"):  # same math, different summation order: bf16 rounding noise only
            m = r[key]
            assert m["ratio"] <= 2e-2 and m["rel_fro"] <= 6e-3, (rank, key, m)
        # v2a merges partial attentions: one lse_merge per bridge layer
        assert r["calls"]["lse_merge"] == 2
        assert r["sp_ok"]
        assert r["reload_ok"], "stale context-parallel weight copies after load_state_dict"
    # every rank ends with the same full-length outputs
    for rank in range(1, world):
        assert torch.equal(results[0]["v2"], results[rank]["v2"]) and torch.equal(results[0]["a2"], results[rank]["a2"])


# 60 tokens -> 30 + 30; 9 tokens -> 5 + 4 (ragged); 4 ranks: 18 tokens -> 5 + 5 + 5 + 3, one video head per rank;
# 6 heads on 2 ranks: an odd head count per rank (like 5 at cp = 8): exchanged head by head, attended as sets [[0, 1], [2]]
@pytest.mark.parametrize("world,grid,heads", [(2, (3, 4, 5), None), (2, (1, 3, 3), None), (2, (2, 3, 3), 4),
                                              (4, (2, 3, 3), 4), (2, (2, 3, 3), 6)])
def test_step_context_parallel_gloo(world, grid, heads):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), grid, heads, results), nprocs=world, join=True)
    _check(results, world, grid, heads)


# The same cases through the peer-memory exchange (copy pushes into the peers' windows + flag words instead of
# all_to_all_single): ragged chunks, several head groups per rank, 4 ranks, and an odd head count per rank attended as
# the sets [[0], [1], [2]] / [[0], [1, 2]] -- results must equal the collective path's, bit for bit on every rank.
@pytest.mark.parametrize("world,grid,heads,set_sizes", [(2, (1, 3, 3), None, None), (4, (2, 3, 3), 4, None),
                                                        (2, (2, 3, 3), 6, (1, 1, 1)), (2, (2, 3, 3), 6, (1, 2))])
def test_step_context_parallel_peer_exchange(world, grid, heads, set_sizes):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), grid, heads, results, "peer", set_sizes), nprocs=world, join=True)
    _check(results, world, grid, heads)
    video_layers = O.TINY_STEP_CFG["visual_layers"]
    for rank in range(world):
        p = results[rank]["peer"]
        assert p["same_as_collective"], f"rank {rank}: peer exchange and all_to_all_single disagree"
        # 4 context-parallel forwards through the windows, one exchange round (epoch) per video self-attention
        assert p["epoch"] == 4 * video_layers and p["waits"] > 0 and p["pushes"] > 0 and p["generation"] >= 1, p
