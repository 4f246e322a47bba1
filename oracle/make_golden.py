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


def build_reference_step(cfg, Pv, Pa, Pb):
    """Reference WanModel / WanAudioModel / bridge holding the oracle's step-level weights, and a pipeline stand-in
    whose two methods are the reference's own source (lifted by ref_loader)."""
    R = ref_loader.load()
    common = dict(text_dim=cfg["text_dim"], freq_dim=cfg["freq_dim"], eps=cfg["eps"], has_image_input=False)
    vis = R.wan_video_dit.WanModel(dim=cfg["visual_dim"], in_dim=cfg["visual_in_dim"], ffn_dim=cfg["visual_ffn"],
                                   out_dim=cfg["visual_out_dim"], patch_size=tuple(cfg["visual_patch"]),
                                   num_heads=cfg["visual_heads"], num_layers=cfg["visual_layers"], **common)
    aud = R.wan_audio_dit.WanAudioModel(dim=cfg["audio_dim"], in_dim=cfg["audio_in_dim"], ffn_dim=cfg["audio_ffn"],
                                        out_dim=cfg["audio_out_dim"], patch_size=list(cfg["audio_patch"]),
                                        num_heads=cfg["audio_heads"], num_layers=cfg["audio_layers"], vae_type="dac",
                                        **common)
    bridge = R.interactionv2.DualTowerConditionalBridge(
        visual_layers=cfg["visual_layers"], audio_layers=cfg["audio_layers"], visual_hidden_dim=cfg["visual_dim"],
        audio_hidden_dim=cfg["audio_dim"], audio_fps=cfg["audio_fps"], head_dim=cfg["head_dim"],
        interaction_strategy=cfg["interaction_strategy"], apply_cross_rope=cfg["apply_cross_rope"])
    vis.load_state_dict(Pv, strict=True)
    aud.load_state_dict(Pa, strict=True)
    bridge.load_state_dict(Pb, strict=True)
    pipe = types.SimpleNamespace(audio_dit=aud, dual_tower_bridge=bridge, _pre_forward=lambda model: None)
    pipe.forward_dual_tower_dit = types.MethodType(R.forward_dual_tower_dit, pipe)
    pipe.inference_single_step = types.MethodType(R.inference_single_step, pipe)
    return R, vis, aud, bridge, pipe


@torch.no_grad()
def run_reference_step(cfg, seed):
    """One MOVA.inference_single_step (pipeline_mova.py:500-609) of the reference in fp32 on CPU, with the
    intermediate embeddings the B200 step memoises."""
    import warnings

    Pv, Pa, Pb, inp = O.make_step_case(cfg, seed)
    R, vis, aud, bridge, pipe = build_reference_step(cfg, Pv, Pa, Pb)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # torch.autocast("cuda") without a GPU only warns and disables itself
        v, a = pipe.inference_single_step(visual_dit=vis, visual_latents=inp["visual_latents"],
                                          audio_latents=inp["audio_latents"], context=inp["context"],
                                          timestep=inp["timestep"], audio_timestep=None, video_fps=cfg["video_fps"])
    out["visual_output"], out["audio_output"] = v, a
    e = R.wan_video_dit.sinusoidal_embedding_1d(cfg["freq_dim"], inp["timestep"])
    out["sinusoidal"] = e
    out["visual_t"] = vis.time_embedding(e)
    out["visual_t_mod"] = vis.time_projection(out["visual_t"]).unflatten(1, (6, cfg["visual_dim"]))
    out["audio_t"] = aud.time_embedding(e)
    out["visual_context"] = vis.text_embedding(inp["context"])
    out["audio_context"] = aud.text_embedding(inp["context"])
    out["visual_tokens"], grid = vis.patchify(inp["visual_latents"])
    out["audio_tokens"], agrid = aud.patchify(inp["audio_latents"], None)
    assert tuple(grid) == tuple(cfg["grid_size"]) and tuple(agrid) == (cfg["audio_len"],)
    hv = vis.head(out["visual_tokens"], out["visual_t"])
    out["visual_head"] = hv
    out["visual_unpatchify"] = vis.unpatchify(hv, grid)
    ha = aud.head(out["audio_tokens"], out["audio_t"])
    out["audio_unpatchify"] = aud.unpatchify(ha, agrid)
    # the full single-tower forwards (wan_video_dit.py:418-473, wan_audio_dit.py:197-252)
    out["video_tower_forward"] = vis(inp["visual_latents"], inp["timestep"], inp["context"])
    out["audio_tower_forward"] = aud(inp["audio_latents"], inp["timestep"], inp["context"])
    keys = {"wan_model": sorted(vis.state_dict().keys()), "wan_audio_model": sorted(aud.state_dict().keys())}
    return {k: t.detach().numpy() for k, t in out.items()}, O.checksum(Pv, Pa, Pb, inp), keys


BRIDGE_BF16_CFG = dict(O.TINY_CFG, grid_size=(9, 2, 2), audio_len=403)  # positions up to 402: the rounding matters


@torch.no_grad()
def run_reference_bridge_bf16(cfg, seed):
    """The reference bridge run the way the reference runs it: ``bridge.to(torch.bfloat16)`` (what
    ``from_pretrained(torch_dtype=bfloat16)`` does, rounding the non-persistent ``inv_freq`` buffer,
    interactionv2.py:21-23) and ``build_aligned_freqs(dtype=bfloat16)`` (pipeline_mova.py:641-648), then one bridge
    layer forward in bf16 on CPU.  Pins ``rope_precision="reference_bf16"`` of the oracle and of the product."""
    Pv, Pa, Pb, inp = O.make_case(cfg, seed)
    R, vis, aud, bridge, pipe = build_reference(cfg, Pv, Pa, Pb)
    bridge = bridge.to(torch.bfloat16)
    assert bridge.rotary.inv_freq.dtype == torch.bfloat16  # the cast reached the buffer
    v_cs, a_cs = bridge.build_aligned_freqs(video_fps=cfg["video_fps"], grid_size=cfg["grid_size"],
                                            audio_steps=cfg["audio_len"], device=torch.device("cpu"),
                                            dtype=torch.bfloat16)
    out = {"cos_v": v_cs[0], "sin_v": v_cs[1], "cos_a": a_cs[0], "sin_a": a_cs[1]}
    xv, xa = inp["visual_x"].to(torch.bfloat16), inp["audio_x"].to(torch.bfloat16)
    bv, ba = bridge(0, xv, xa, x_freqs=v_cs, y_freqs=a_cs, condition_scale=1.0, video_grid_size=cfg["grid_size"])
    out["bridge0_visual"], out["bridge0_audio"] = bv, ba
    # the same layer with the reference in fp32 (exact tables): how far the two precisions are apart
    R2, _, _, bridge32, _ = build_reference(cfg, Pv, Pa, Pb)
    v32, a32 = bridge32.build_aligned_freqs(video_fps=cfg["video_fps"], grid_size=cfg["grid_size"],
                                            audio_steps=cfg["audio_len"], device=torch.device("cpu"), dtype=torch.float32)
    bv32, ba32 = bridge32(0, inp["visual_x"], inp["audio_x"], x_freqs=v32, y_freqs=a32, condition_scale=1.0,
                          video_grid_size=cfg["grid_size"])
    out["bridge0_visual_fp32"], out["bridge0_audio_fp32"] = bv32, ba32
    return {k: v.detach().float().numpy() for k, v in out.items()}, O.checksum(Pv, Pa, Pb, inp)


def sample_indices(numel: int, n: int, seed: int = 0):
    """Fixed pseudo-random flat indices into a tensor of ``numel`` elements (shared by the generator and the tests)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (n,), generator=g)


@torch.no_grad()
def run_reference_reduced(cfg, seed):
    """BASELINE.json configs[0]: the reference's dual-tower forward at MOVA-360p WIDTHS and reduced depth (2 + 2 blocks,
    2 bridge layers, 352x640x17-frame clip: L_v = 4400, L_a = 36, 512 text tokens), fp32 on CPU.  The outputs are 90 MB,
    so the fixture keeps 8192 + 2048 sampled elements and whole-tensor statistics."""
    Pv, Pa, Pb, inp = O.make_case(cfg, seed)
    R, vis, aud, bridge, pipe = build_reference(cfg, Pv, Pa, Pb)
    fv, fa = R.forward_dual_tower_dit(
        pipe, visual_dit=vis, visual_x=inp["visual_x"], audio_x=inp["audio_x"], visual_context=inp["visual_context"],
        audio_context=inp["audio_context"], visual_t_mod=inp["visual_t_mod"], audio_t_mod=inp["audio_t_mod"],
        visual_freqs=inp["visual_freqs"], audio_freqs=inp["audio_freqs"], grid_size=cfg["grid_size"],
        video_fps=cfg["video_fps"])
    out = {}
    for name, t, n in (("visual", fv, 8192), ("audio", fa, 2048)):
        idx = sample_indices(t.numel(), n, seed=11)
        out[f"{name}_samples"] = t.reshape(-1)[idx]
        out[f"{name}_stats"] = torch.tensor([t.mean().item(), t.std().item(), t.abs().max().item(), t.norm().item()])
        out[f"{name}_delta_samples"] = (t - inp[f"{name if name == 'audio' else 'visual'}_x"]).reshape(-1)[idx]
    return {k: v.numpy() for k, v in out.items()}, O.checksum(Pv, Pa, Pb, inp)


def write_bridge_bf16(out_dir):
    arrays, csum = run_reference_bridge_bf16(BRIDGE_BF16_CFG, 31)
    meta = dict(cfg=BRIDGE_BF16_CFG, seed=31, checksum=csum, torch=torch.__version__,
                source="reference DualTowerConditionalBridge after .to(torch.bfloat16): build_aligned_freqs(dtype=bf16) "
                       "tables and one bridge layer forward in bf16 on CPU (+ the same layer in fp32), by "
                       "oracle/make_golden.py")
    np.savez_compressed(os.path.join(out_dir, "bridge_rope_bf16.npz"), **arrays)
    with open(os.path.join(out_dir, "bridge_rope_bf16.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("bridge_rope_bf16", {k: v.shape for k, v in arrays.items()}, "checksum", csum)


def main():
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "bridge_bf16":  # only the newest fixture (the others are unchanged)
        write_bridge_bf16(out_dir)
        return
    write_bridge_bf16(out_dir)
    for name, cfg, seed in (("tiny_dual_tower", O.TINY_CFG, 1234),):
        arrays, csum, keys = run_reference(cfg, seed)
        meta = dict(cfg=cfg, seed=seed, checksum=csum, torch=torch.__version__, reference_state_dict_keys=keys,
                    source="reference modules from /root/reference run in fp32 on CPU by oracle/make_golden.py")
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        print(name, {k: v.shape for k, v in arrays.items()}, "checksum", csum)
    if os.environ.get("MOVA_GOLDEN_REDUCED", "1") == "1":  # ~1 minute and ~6 GB of RAM on 8 cores
        arrays, csum = run_reference_reduced(O.REDUCED_360P_CFG, 2024)
        meta = dict(cfg=O.REDUCED_360P_CFG, seed=2024, checksum=csum, torch=torch.__version__, sample_seed=11,
                    source="reference DiTBlocks / bridge + MOVA.forward_dual_tower_dit at MOVA-360p widths (BASELINE "
                           "configs[0]) run in fp32 on CPU by oracle/make_golden.py; sampled elements + statistics")
        np.savez_compressed(os.path.join(out_dir, "reduced_360p_samples.npz"), **arrays)
        with open(os.path.join(out_dir, "reduced_360p_samples.json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        print("reduced_360p_samples", {k: v.shape for k, v in arrays.items()}, "checksum", csum)
    for name, cfg, seed in (("tiny_step", O.TINY_STEP_CFG, 1234),):
        arrays, csum, keys = run_reference_step(cfg, seed)
        meta = dict(cfg=cfg, seed=seed, checksum=csum, torch=torch.__version__, reference_state_dict_keys=keys,
                    source="reference WanModel / WanAudioModel / bridge + MOVA.inference_single_step (source lifted from "
                           "pipeline_mova.py:500-609) run in fp32 on CPU by oracle/make_golden.py")
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        print(name, {k: v.shape for k, v in arrays.items()}, "checksum", csum)


if __name__ == "__main__":
    main()
