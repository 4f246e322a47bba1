"""TEST INFRASTRUCTURE ONLY -- generates the committed golden vectors under tests/golden/ by running the
UNMODIFIED reference modules (imported from /root/reference through oracle/ref_loader.py) in fp32 on CPU.

    TORCHDYNAMO_DISABLE=1 python oracle/make_golden.py

The reference ships no tests or golden vectors, so these outputs of the reference's own code are what pins the
oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.  Weights and inputs are NOT stored: they are
regenerated from the seed by ``mova_oracle.make_case`` and guarded by a checksum stored beside the outputs.
"""
from __future__ import annotations

import json
import os
import sys
import types

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import mova_oracle as O  # noqa: E402
import ref_loader  # noqa: E402


def build_reference(cfg, Pv, Pa, Pb):
    """Reference modules holding the oracle's weights (state-dict keys are the reference's own)."""
    R = ref_loader.load()
    DiTBlock = R.wan_video_dit.DiTBlock
    vis = torch.nn.Module()
    vis.blocks = torch.nn.ModuleList([DiTBlock(False, cfg["visual_dim"], cfg["visual_heads"], cfg["visual_ffn"], cfg["eps"])
                                      for _ in range(cfg["visual_layers"])])
    aud = torch.nn.Module()
    aud.blocks = torch.nn.ModuleList([DiTBlock(False, cfg["audio_dim"], cfg["audio_heads"], cfg["audio_ffn"], cfg["eps"])
                                      for _ in range(cfg["audio_layers"])])
    bridge = R.interactionv2.DualTowerConditionalBridge(
        visual_layers=cfg["visual_layers"], audio_layers=cfg["audio_layers"], visual_hidden_dim=cfg["visual_dim"],
        audio_hidden_dim=cfg["audio_dim"], audio_fps=cfg["audio_fps"], head_dim=cfg["head_dim"],
        interaction_strategy=cfg["interaction_strategy"], apply_cross_rope=cfg["apply_cross_rope"])
    missing, unexpected = vis.load_state_dict(Pv, strict=True), None
    aud.load_state_dict(Pa, strict=True)
    bridge.load_state_dict(Pb, strict=True)
    del missing, unexpected
    pipe = types.SimpleNamespace(audio_dit=aud, dual_tower_bridge=bridge)
    return R, vis, aud, bridge, pipe


@torch.no_grad()
def run_reference(cfg, seed):
    Pv, Pa, Pb, inp = O.make_case(cfg, seed)
    R, vis, aud, bridge, pipe = build_reference(cfg, Pv, Pa, Pb)
    out = {}
    # single modules
    out["video_block0"] = vis.blocks[0](inp["visual_x"], inp["visual_context"], inp["visual_t_mod"], inp["visual_freqs"])
    out["audio_block0"] = aud.blocks[0](inp["audio_x"], inp["audio_context"], inp["audio_t_mod"], inp["audio_freqs"])
    v_cs, a_cs = bridge.build_aligned_freqs(video_fps=cfg["video_fps"], grid_size=cfg["grid_size"],
                                            audio_steps=cfg["audio_len"], device=torch.device("cpu"), dtype=torch.float32)
    out["cos_v"], out["sin_v"], out["cos_a"], out["sin_a"] = v_cs[0], v_cs[1], a_cs[0], a_cs[1]
    bv, ba = bridge(0, inp["visual_x"], inp["audio_x"], x_freqs=v_cs, y_freqs=a_cs, condition_scale=1.0,
                    video_grid_size=cfg["grid_size"])
    out["bridge0_visual"], out["bridge0_audio"] = bv, ba
    # the whole path, the reference's own loop (pipeline_mova.py:612-711)
    fv, fa = R.forward_dual_tower_dit(
        pipe, visual_dit=vis, visual_x=inp["visual_x"], audio_x=inp["audio_x"], visual_context=inp["visual_context"],
        audio_context=inp["audio_context"], visual_t_mod=inp["visual_t_mod"], audio_t_mod=inp["audio_t_mod"],
        visual_freqs=inp["visual_freqs"], audio_freqs=inp["audio_freqs"], grid_size=cfg["grid_size"],
        video_fps=cfg["video_fps"])
    out["final_visual"], out["final_audio"] = fv, fa
    # CP helpers (functional.py:55-111) on the audio tokens, 4 ranks: ragged last chunk
    for r in range(4):
        out[f"sp_split_r{r}"] = R.functional._sp_split_tensor(inp["audio_x"], sp_size=4, sp_rank=r)[0]
    keys = {"dit_block": sorted(vis.blocks[0].state_dict().keys()), "bridge": sorted(bridge.state_dict().keys())}
    return {k: v.detach().numpy() for k, v in out.items()}, O.checksum(Pv, Pa, Pb, inp), keys


def main():
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, cfg, seed in (("tiny_dual_tower", O.TINY_CFG, 1234),):
        arrays, csum, keys = run_reference(cfg, seed)
        meta = dict(cfg=cfg, seed=seed, checksum=csum, torch=torch.__version__, reference_state_dict_keys=keys,
                    source="reference modules from /root/reference run in fp32 on CPU by oracle/make_golden.py")
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        print(name, {k: v.shape for k, v in arrays.items()}, "checksum", csum)


if __name__ == "__main__":
    main()
