"""Shared helpers of the parity tests: build the B200 modules from an oracle weight dict, error metrics."""
import types

import torch


def bf16_round(d):
    """Weights / inputs as the CUDA path sees them (bf16), handed to the fp32 oracle as exact fp32 values."""
    return {k: (v.to(torch.bfloat16).to(torch.float32) if v.is_floating_point() else v) for k, v in d.items()}


def sub_state(P, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in P.items() if k.startswith(prefix)}


def build_towers(cfg, Pv, Pa, Pb, device="cuda"):
    """dualforce_b200 modules holding the given weights in bf16 on ``device`` + a pipeline-like namespace."""
    import dualforce_b200 as B

    vis = torch.nn.Module()
    vis.blocks = torch.nn.ModuleList([B.DiTBlock(False, cfg["visual_dim"], cfg["visual_heads"], cfg["visual_ffn"], cfg["eps"])
                                      for _ in range(cfg["visual_layers"])])
    aud = torch.nn.Module()
    aud.blocks = torch.nn.ModuleList([B.DiTBlock(False, cfg["audio_dim"], cfg["audio_heads"], cfg["audio_ffn"], cfg["eps"])
                                      for _ in range(cfg["audio_layers"])])
    bridge = B.DualTowerConditionalBridge(
        visual_layers=cfg["visual_layers"], audio_layers=cfg["audio_layers"], visual_hidden_dim=cfg["visual_dim"],
        audio_hidden_dim=cfg["audio_dim"], audio_fps=cfg["audio_fps"], head_dim=cfg["head_dim"],
        interaction_strategy=cfg["interaction_strategy"], apply_cross_rope=cfg["apply_cross_rope"])
    vis.load_state_dict(Pv, strict=True)
    aud.load_state_dict(Pa, strict=True)
    bridge.load_state_dict(Pb, strict=True)
    for m in (vis, aud, bridge):
        m.to(device=device, dtype=torch.bfloat16)
    pipe = types.SimpleNamespace(audio_dit=aud, dual_tower_bridge=bridge, video_dit=vis, video_dit_2=None)
    pipe.forward_dual_tower_dit = types.MethodType(B.forward_dual_tower_dit, pipe)
    return vis, aud, bridge, pipe


def build_step_towers(cfg, Pv, Pa, Pb, device="cuda"):
    """dualforce_b200 WanModel / WanAudioModel / bridge twins holding step-level weights (oracle.make_step_case) in
    bf16 on ``device`` + a pipeline-like namespace with both drop-in methods bound."""
    import dualforce_b200 as B

    common = dict(text_dim=cfg["text_dim"], freq_dim=cfg["freq_dim"], eps=cfg["eps"], has_image_input=False)
    vis = B.WanModel(dim=cfg["visual_dim"], in_dim=cfg["visual_in_dim"], ffn_dim=cfg["visual_ffn"],
                     out_dim=cfg["visual_out_dim"], patch_size=tuple(cfg["visual_patch"]), num_heads=cfg["visual_heads"],
                     num_layers=cfg["visual_layers"], **common)
    aud = B.WanAudioModel(dim=cfg["audio_dim"], in_dim=cfg["audio_in_dim"], ffn_dim=cfg["audio_ffn"],
                          out_dim=cfg["audio_out_dim"], patch_size=list(cfg["audio_patch"]), num_heads=cfg["audio_heads"],
                          num_layers=cfg["audio_layers"], vae_type="dac", **common)
    bridge = B.DualTowerConditionalBridge(
        visual_layers=cfg["visual_layers"], audio_layers=cfg["audio_layers"], visual_hidden_dim=cfg["visual_dim"],
        audio_hidden_dim=cfg["audio_dim"], audio_fps=cfg["audio_fps"], head_dim=cfg["head_dim"],
        interaction_strategy=cfg["interaction_strategy"], apply_cross_rope=cfg["apply_cross_rope"])
    vis.load_state_dict(Pv, strict=True)
    aud.load_state_dict(Pa, strict=True)
    bridge.load_state_dict(Pb, strict=True)
    for m in (vis, aud, bridge):
        m.to(device=device, dtype=torch.bfloat16)
    pipe = types.SimpleNamespace(audio_dit=aud, dual_tower_bridge=bridge, video_dit=vis, video_dit_2=None)
    pipe.forward_dual_tower_dit = types.MethodType(B.forward_dual_tower_dit, pipe)
    pipe.inference_single_step = types.MethodType(B.inference_single_step, pipe)
    return vis, aud, bridge, pipe


def to_dev(inp, device="cuda"):
    out = {}
    for k, v in inp.items():
        if torch.is_complex(v):
            out[k] = v.to(device)
        else:
            out[k] = v.to(device=device, dtype=torch.bfloat16)
    return out


def metrics(got, ref):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    err = (got - ref).abs().max().item()
    amax = ref.abs().max().item()
    fro = ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    return dict(max_abs=err, abs_max=amax, ratio=err / max(amax, 1e-30), rel_fro=fro, cos=cos,
                finite=bool(torch.isfinite(got).all()))


# bf16 tolerance of the path: about 2x the reference's own bf16-vs-fp32 noise
TOL_RATIO = 1.5e-2     # max abs err / abs-max of the reference output
TOL_FRO = 6e-3         # relative Frobenius error
TOL_DELTA_COS = 0.995  # cosine of the residual delta  block(x) - x


def assert_close(got, ref, name, ratio=TOL_RATIO, fro=TOL_FRO):
    m = metrics(got, ref)
    assert m["finite"], f"{name}: non-finite output"
    assert m["ratio"] <= ratio and m["rel_fro"] <= fro, f"{name}: {m}"
    return m
