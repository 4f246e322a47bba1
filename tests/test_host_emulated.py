"""Host logic of the product (module twins, weight packing, stride plumbing, the dual-tower loop, the step wrapper and
its memo caches) run on CPU through tests/emulated_ops.py -- a test-only stand-in for libmova_b200.so -- and checked
against the oracle and the reference's golden outputs.  The CUDA kernels themselves are covered by the -m gpu tests."""
import json
import os
import types

import numpy as np
import pytest
import torch

import emulated_ops
import mova_oracle as O
import ref_loader
from test_oracle_step_golden import load_step_case
from util import assert_close, bf16_round, build_step_towers, build_towers, to_dev

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture()
def emu(monkeypatch):
    import dualforce_b200 as B

    B.rope.clear_cache()
    return emulated_ops.install(monkeypatch)


def test_dual_tower_forward_host_logic_vs_oracle_and_golden(emu):
    with open(os.path.join(GOLDEN, "tiny_dual_tower.json")) as f:
        meta = json.load(f)
    cfg = meta["cfg"]
    cfg["grid_size"] = tuple(cfg["grid_size"])
    gold = np.load(os.path.join(GOLDEN, "tiny_dual_tower.npz"))
    Pv, Pa, Pb, inp = O.make_case(cfg, meta["seed"])
    Pv, Pa, Pb, inp = bf16_round(Pv), bf16_round(Pa), bf16_round(Pb), bf16_round(inp)
    vis, aud, bridge, pipe = build_towers(cfg, Pv, Pa, Pb, device="cpu")
    d = to_dev(inp, device="cpu")
    fv, fa = pipe.forward_dual_tower_dit(vis, d["visual_x"], d["audio_x"], d["visual_context"], d["audio_context"],
                                         d["visual_t_mod"], d["audio_t_mod"], d["visual_freqs"], d["audio_freqs"],
                                         cfg["grid_size"], cfg["video_fps"])
    rv, ra = O.forward_dual_tower_dit(Pv, Pa, Pb, cfg, inp["visual_x"], inp["audio_x"], inp["visual_context"],
                                      inp["audio_context"], inp["visual_t_mod"], inp["audio_t_mod"], inp["visual_freqs"],
                                      inp["audio_freqs"], cfg["grid_size"], cfg["video_fps"])
    assert_close(fv, rv, "visual vs oracle", ratio=3e-2, fro=1.2e-2)
    assert_close(fa, ra, "audio vs oracle", ratio=3e-2, fro=1.2e-2)
    assert_close(fv, torch.from_numpy(gold["final_visual"]), "visual vs reference golden", ratio=3e-2, fro=1.2e-2)
    assert_close(fa, torch.from_numpy(gold["final_audio"]), "audio vs reference golden", ratio=3e-2, fro=1.2e-2)
    # 17 launches per DiTBlock, 7 + 7 per bridge layer: nothing else may reach the device
    n_blocks, n_bridge = cfg["visual_layers"] + cfg["audio_layers"], min(cfg["visual_layers"], cfg["audio_layers"])
    assert sum(emu.values()) == 17 * n_blocks + 14 * n_bridge, emu
    # inputs must not be mutated
    assert torch.equal(d["visual_x"].float(), inp["visual_x"]) and torch.equal(d["audio_x"].float(), inp["audio_x"])


@pytest.fixture()
def step_case():
    cfg, Pv, Pa, Pb, inp, gold, meta = load_step_case()
    Pv, Pa, Pb = bf16_round(Pv), bf16_round(Pa), bf16_round(Pb)
    inp = dict(inp, context=inp["context"].to(torch.bfloat16).float())
    return cfg, Pv, Pa, Pb, inp, gold, meta


def test_patchify_index_math_is_exact(emu):
    """The kernels' flat-index arithmetic (transcribed in emulated_ops) against the oracle's reshape/permute form."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 6, 4, 9, generator=g).to(torch.bfloat16)  # C, F, H, W with patch (2, 2, 3)
    cols = emulated_ops.patchify(x, (2, 2, 3))
    eye = {"patch_embedding.weight": torch.eye(60).reshape(60, 5, 2, 2, 3), "patch_embedding.bias": torch.zeros(60)}
    ref, grid = O.patchify(eye, x.float()[None], (2, 2, 3))
    assert grid == (3, 2, 3) and torch.equal(cols.float(), ref[0])
    y = torch.randn(3 * 2 * 3, 2 * 2 * 3 * 7, generator=g).to(torch.bfloat16)
    out = emulated_ops.unpatchify(y, grid, (2, 2, 3), 7)
    assert torch.equal(out.float(), O.unpatchify(y.float()[None], grid, (2, 2, 3))[0])
    # 1-D (audio) form, from a strided view
    buf = torch.randn(11, 40, generator=g).to(torch.bfloat16)
    out = emulated_ops.unpatchify(buf[:, 8:32], (11,), (3,), 8)
    assert torch.equal(out.float(), O.unpatchify(buf[:, 8:32].float()[None], (11,), (3,))[0])


def test_step_pieces_vs_reference_golden(emu, step_case):
    import dualforce_b200 as B
    from dualforce_b200 import step

    cfg, Pv, Pa, Pb, inp, gold, _ = step_case
    vis, aud, bridge, pipe = build_step_towers(cfg, Pv, Pa, Pb, device="cpu")
    t, t_mod = step.embed_time(vis, inp["timestep"])
    assert t.dtype == torch.bfloat16 and t.shape == (1, cfg["visual_dim"]) and t_mod.shape == (1, 6, cfg["visual_dim"])
    assert_close(t, gold["visual_t"], "t", ratio=8e-3, fro=6e-3)
    assert_close(t_mod, gold["visual_t_mod"], "t_mod", ratio=8e-3, fro=6e-3)
    ctx = step.embed_text(vis, inp["context"].to(torch.bfloat16))
    assert_close(ctx, gold["visual_context"], "text embedding", ratio=1.5e-2, fro=8e-3)
    tok, grid = step.patchify(vis, inp["visual_latents"])
    assert tuple(grid) == cfg["grid_size"]
    assert_close(tok, gold["visual_tokens"], "video patchify", ratio=1.5e-2, fro=8e-3)
    tok_a, (f,) = step.patchify(aud, inp["audio_latents"])
    assert f == cfg["audio_len"]
    assert_close(tok_a, gold["audio_tokens"], "audio patchify", ratio=1.5e-2, fro=8e-3)
    out = step.head_unpatchify(vis, gold["visual_tokens"].to(torch.bfloat16), gold["visual_t"].to(torch.bfloat16), grid)
    assert_close(out, gold["visual_unpatchify"], "video head + unpatchify", ratio=1.5e-2, fro=8e-3)
    out_a = step.head_unpatchify(aud, gold["audio_tokens"].to(torch.bfloat16), gold["audio_t"].to(torch.bfloat16), (f,))
    assert_close(out_a, gold["audio_unpatchify"], "audio head + unpatchify", ratio=1.5e-2, fro=8e-3)
    # RoPE tables as the reference pipeline assembles them
    assert torch.equal(step.token_freqs(vis, grid, "cpu"), O.video_freqs(cfg["head_dim"], grid))
    assert torch.equal(step.token_freqs(aud, (f,), "cpu"), O.audio_freqs(cfg["head_dim"], f))
    # single-tower forwards (wan_video_dit.py:418-473)
    assert_close(vis(inp["visual_latents"], inp["timestep"], inp["context"].to(torch.bfloat16)),
                 gold["video_tower_forward"], "WanModel.forward", ratio=3e-2, fro=1.5e-2)
    assert_close(aud(inp["audio_latents"], inp["timestep"], inp["context"].to(torch.bfloat16)),
                 gold["audio_tower_forward"], "WanAudioModel.forward", ratio=3e-2, fro=1.5e-2)
    assert B.WanModel is step.WanModel


def test_inference_single_step_vs_oracle_and_golden(emu, step_case):
    cfg, Pv, Pa, Pb, inp, gold, _ = step_case
    vis, aud, bridge, pipe = build_step_towers(cfg, Pv, Pa, Pb, device="cpu")
    ctx = inp["context"].to(torch.bfloat16)
    v, a = pipe.inference_single_step(visual_dit=vis, visual_latents=inp["visual_latents"],
                                      audio_latents=inp["audio_latents"], context=ctx, timestep=inp["timestep"],
                                      audio_timestep=None, video_fps=cfg["video_fps"])
    assert v.dtype == torch.bfloat16 and v.shape == gold["visual_output"].shape and a.shape == gold["audio_output"].shape
    rv, ra = O.inference_single_step(Pv, Pa, Pb, cfg, inp["visual_latents"], inp["audio_latents"], inp["context"],
                                     inp["timestep"])
    assert_close(v, rv, "step visual vs oracle", ratio=3e-2, fro=1.5e-2)
    assert_close(a, ra, "step audio vs oracle", ratio=3e-2, fro=1.5e-2)
    assert_close(v, gold["visual_output"], "step visual vs reference golden", ratio=3e-2, fro=1.5e-2)
    assert_close(a, gold["audio_output"], "step audio vs reference golden", ratio=3e-2, fro=1.5e-2)


def test_step_memoises_prompt_and_timestep_work(emu, step_case):
    """Second denoising step with the same two prompts: text embeddings and every layer's text k/v come from the
    memo; within a step the negative call reuses the positive call's time embedding."""
    cfg, Pv, Pa, Pb, inp, gold, _ = step_case
    vis, aud, bridge, pipe = build_step_towers(cfg, Pv, Pa, Pb, device="cpu")
    pos = inp["context"].to(torch.bfloat16)
    neg = torch.zeros_like(pos)

    def one_step(ts):
        outs = []
        for ctx in (pos, neg):
            before = dict(emu)
            outs.append(pipe.inference_single_step(visual_dit=vis, visual_latents=inp["visual_latents"],
                                                   audio_latents=inp["audio_latents"], context=ctx, timestep=ts,
                                                   audio_timestep=None, video_fps=cfg["video_fps"]))
            outs.append({k: emu.get(k, 0) - before.get(k, 0) for k in emu})
        return outs

    n_blocks = cfg["visual_layers"] + cfg["audio_layers"]
    (v1, a1), c_pos1, (v1n, a1n), c_neg1 = one_step(torch.tensor([900.0]))
    (v2, a2), c_pos2, (v2n, a2n), c_neg2 = one_step(torch.tensor([880.0]))
    assert c_pos1["gemv_f32"] == 6 and c_pos1["sinusoidal_embedding"] == 2  # both towers
    assert c_neg1.get("gemv_f32", 0) == 0 and c_neg1.get("sinusoidal_embedding", 0) == 0  # same timestep tensor
    assert c_pos2["gemv_f32"] == 6  # new timestep
    # text embedding: 2 GEMMs per tower; text k/v: one GEMM + one RMSNorm per block -- all skipped on step 2
    assert c_pos1["linear"] - c_pos2["linear"] == 4 + n_blocks
    assert c_neg1["linear"] - c_neg2["linear"] == 4 + n_blocks
    assert c_pos1["rmsnorm_rope_"] - c_pos2["rmsnorm_rope_"] == n_blocks
    # and the memo changes nothing: a fresh (uncached) evaluation of step 2 gives bit-identical outputs
    from dualforce_b200 import step

    step.clear_step_caches(vis, aud)
    (v3, a3), _, (v3n, a3n), _ = one_step(torch.tensor([880.0]))
    assert torch.equal(v2, v3) and torch.equal(a2, a3) and torch.equal(v2n, v3n) and torch.equal(a2n, a3n)
    assert not torch.equal(v1, v2) and not torch.equal(v2, v2n)
    # an in-place edit of the prompt bumps its version: no stale hit
    pos.mul_(0.5)
    (v4, _), c4, *_ = one_step(torch.tensor([880.0]))
    assert c4["linear"] == c_pos1["linear"] and not torch.equal(v4, v2)


def test_weight_reload_invalidates_every_memo(emu, step_case):
    """load_state_dict() after a step keeps every data_ptr (the packed buffers are updated in place) but bumps the
    parameter versions: time / text embeddings, per-layer text k / v and the packed weights must all follow.  A step
    on reloaded towers is bit-identical to the same step on towers built from the new weights."""
    cfg, Pv, Pa, Pb, inp, gold, _ = step_case
    vis, aud, bridge, pipe = build_step_towers(cfg, Pv, Pa, Pb, device="cpu")
    ctx = inp["context"].to(torch.bfloat16)
    kw = dict(visual_latents=inp["visual_latents"], audio_latents=inp["audio_latents"], context=ctx,
              timestep=inp["timestep"], audio_timestep=None, video_fps=cfg["video_fps"])
    v_old, a_old = pipe.inference_single_step(visual_dit=vis, **kw)
    Qv, Qa, Qb, _ = O.make_step_case(cfg, 4242)
    Qv, Qa, Qb = bf16_round(Qv), bf16_round(Qa), bf16_round(Qb)
    ptr = vis.blocks[0].self_attn.q.weight.data_ptr()
    vis.load_state_dict({k: v.to(torch.bfloat16) for k, v in Qv.items()})
    aud.load_state_dict({k: v.to(torch.bfloat16) for k, v in Qa.items()})
    bridge.load_state_dict({k: v.to(torch.bfloat16) for k, v in Qb.items()})
    assert vis.blocks[0].self_attn.q.weight.data_ptr() == ptr  # in place: only the version tells
    v_new, a_new = pipe.inference_single_step(visual_dit=vis, **kw)
    vis2, aud2, bridge2, pipe2 = build_step_towers(cfg, Qv, Qa, Qb, device="cpu")
    v_ref, a_ref = pipe2.inference_single_step(visual_dit=vis2, **kw)
    assert torch.equal(v_new, v_ref) and torch.equal(a_new, a_ref)
    assert not torch.equal(v_new, v_old)


def test_twin_state_dict_keys_equal_the_reference(step_case):
    cfg, Pv, Pa, Pb, inp, gold, meta = step_case
    vis, aud, bridge, pipe = build_step_towers(cfg, Pv, Pa, Pb, device="cpu")
    assert sorted(vis.state_dict().keys()) == meta["reference_state_dict_keys"]["wan_model"]
    assert sorted(aud.state_dict().keys()) == meta["reference_state_dict_keys"]["wan_audio_model"]


def test_unsupported_step_options_fail_loudly(emu, step_case):
    import dualforce_b200 as B

    cfg, Pv, Pa, Pb, inp, gold, _ = step_case
    with pytest.raises(NotImplementedError):
        B.WanModel(dim=256, in_dim=36, ffn_dim=512, out_dim=16, text_dim=64, freq_dim=256, eps=1e-6,
                   patch_size=(1, 2, 2), num_heads=2, num_layers=1, has_image_input=True)
    vis, aud, bridge, pipe = build_step_towers(cfg, Pv, Pa, Pb, device="cpu")
    with pytest.raises(ValueError):  # three prompts against two latents: not a broadcastable batch
        pipe.inference_single_step(visual_dit=vis, visual_latents=inp["visual_latents"].repeat(2, 1, 1, 1, 1),
                                   audio_latents=inp["audio_latents"],
                                   context=inp["context"].to(torch.bfloat16).repeat(3, 1, 1),
                                   timestep=inp["timestep"], audio_timestep=None, video_fps=24.0)
    with pytest.raises(NotImplementedError):  # per-token time embedding
        vis.head(torch.zeros(1, 4, 256, dtype=torch.bfloat16), torch.zeros(1, 4, 256, dtype=torch.bfloat16))


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_swap_modules_on_reference_pipeline_runs_the_reference_loop_shape(emu, step_case):
    """dualforce_b200.install's host half on REAL reference towers: the swapped pipeline (reference WanModel
    containers, B200 blocks / heads / bridge sharing their Parameters) reproduces the reference's own step."""
    import make_golden
    from dualforce_b200 import pipeline as pl
    from dualforce_b200 import step

    cfg, Pv, Pa, Pb, inp, gold, _ = step_case
    R, vis, aud, bridge, ref_pipe = make_golden.build_reference_step(cfg, Pv, Pa, Pb)
    for m in (vis, aud, bridge):
        m.to(torch.bfloat16)
    pipe = types.SimpleNamespace(video_dit=vis, video_dit_2=None, audio_dit=aud, dual_tower_bridge=bridge,
                                 _pre_forward=lambda m: None)
    n = pl.swap_modules(pipe)
    n_bridge = 2 * min(cfg["visual_layers"], cfg["audio_layers"])
    assert n == cfg["visual_layers"] + cfg["audio_layers"] + n_bridge + 2
    assert isinstance(vis.head, step.Head) and isinstance(pipe.dual_tower_bridge, pl.DualTowerConditionalBridge)
    assert pl.swap_modules(pipe) == 0  # idempotent
    # the script's own `pipe.replace_attention(...)` call (inference_single.py:115) keeps its return contract and
    # leaves the B200 attention processors in place
    sites = cfg["visual_layers"] + cfg["audio_layers"] + n_bridge
    assert pipe.replace_attention(attn_type="fa") == sites
    assert type(vis.blocks[0].self_attn.attn).__module__.startswith("dualforce_b200")
    v, a = pipe.inference_single_step(visual_dit=vis, visual_latents=inp["visual_latents"],
                                      audio_latents=inp["audio_latents"], context=inp["context"].to(torch.bfloat16),
                                      timestep=inp["timestep"], audio_timestep=None, video_fps=cfg["video_fps"])
    assert_close(v, gold["visual_output"], "swapped reference pipeline, visual", ratio=3e-2, fro=1.5e-2)
    assert_close(a, gold["audio_output"], "swapped reference pipeline, audio", ratio=3e-2, fro=1.5e-2)


def test_bench_model_builder_composes_with_the_step(emu):
    """bench.py's model builder (tower twins with both video experts, bridge, bound drop-in methods) runs the step and
    the fused CFG + Euler update the timed region is made of (host logic only here; tests/dryrun_bench.py runs the
    whole arm)."""
    import bench
    from dualforce_b200 import step

    cfg = dict(O.TINY_CFG, grid_size=(2, 2, 3), audio_len=9)
    pipe = bench.build_model(cfg, torch.device("cpu"), experts=2)
    assert len(pipe.video_dit.blocks) == cfg["visual_layers"] and len(pipe.audio_dit.blocks) == cfg["audio_layers"]
    assert pipe.video_dit_2 is not None and pipe.video_dit_2 is not pipe.video_dit
    assert pipe.video_dit.patch_embedding.weight.dtype == torch.bfloat16
    S = bench.STEP_360P
    host = bench.host_step_inputs(cfg, pin=False)
    assert host["latents"].shape == (1, 16, 2, 4, 6) and host["condition"].shape == (1, 20, 2, 4, 6)
    x_in = torch.cat([host["latents"], host["condition"]], dim=1)
    kw = dict(visual_dit=pipe.video_dit, visual_latents=x_in, audio_latents=host["audio_latents"],
              timestep=torch.tensor([900.0]), audio_timestep=None, video_fps=24.0)
    before = sum(emu.values())
    pv, pa = pipe.inference_single_step(context=host["context_pos"], **kw)
    per_forward_first = sum(emu.values()) - before
    nv, na = pipe.inference_single_step(context=host["context_neg"], **kw)
    assert pv.shape == (1, S["visual_out_dim"], 2, 4, 6) and pa.shape == (1, S["audio_out_dim"], 9)
    assert torch.isfinite(pv.float()).all() and torch.isfinite(pa.float()).all()
    ts, sig = bench.flow_match_schedule(50)
    assert abs(float(ts[0]) - 1000.0) < 1e-3 and float(sig[-1]) > 0 and all(sig[i] > sig[i + 1] for i in range(49))
    lat_next = step.guided_update(pv, nv, host["latents"], 5.0, float(sig[0]), float(sig[1]))
    want = host["latents"] + (nv.float() + 5.0 * (pv.float() - nv.float())) * (float(sig[1]) - float(sig[0]))
    assert torch.allclose(lat_next, want, rtol=1e-5, atol=1e-5)
    n_blocks = cfg["visual_layers"] + cfg["audio_layers"]
    # 17 launches per block + 14 per bridge layer; per forward: 2 patchify + 2 patch GEMMs + 2 x (add, LN, GEMM,
    # unpatchify) heads; first forward of a step: 8 time-embedding kernels + 4 text-embedding GEMMs
    assert per_forward_first == 17 * n_blocks + 14 * min(cfg["visual_layers"], cfg["audio_layers"]) + 4 + 8 + 8 + 4


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_lora_wrapped_reference_block_is_merged_at_install(emu):
    """a reference block whose q/k/v/o were wrapped by the reference's own inject_lora_to_model
    (engine/trainer/accelerate/lora_utils.py) is swapped with the adapters folded into the weights; the result
    matches the reference's un-merged LoRA forward."""
    import importlib.util
    import sys

    import dualforce_b200 as B

    R = ref_loader.load()
    sys.modules["diffusers"].DiffusionPipeline = object  # lora_utils.py imports the name only
    spec = importlib.util.spec_from_file_location(
        "ref_lora_utils", os.path.join(ref_loader.REFERENCE_ROOT, "mova/engine/trainer/accelerate/lora_utils.py"))
    lora_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lora_utils)

    cfg = O.TINY_CFG
    Pv, Pa, Pb, inp = O.make_case(cfg, 5)
    Pv, inp = bf16_round(Pv), bf16_round(inp)
    blk = R.wan_video_dit.DiTBlock(False, cfg["visual_dim"], cfg["visual_heads"], cfg["visual_ffn"], cfg["eps"])
    blk.load_state_dict({k[len("blocks.0."):]: v for k, v in Pv.items() if k.startswith("blocks.0.")})
    layers = lora_utils.inject_lora_to_model(blk, rank=4, alpha=8.0, target_modules=["q", "k", "v", "o"])
    assert len(layers) == 8  # q, k, v, o of self_attn and cross_attn
    g = torch.Generator().manual_seed(9)
    for l in layers.values():
        l.lora_B.weight.data = (torch.randn(l.lora_B.weight.shape, generator=g) * 0.2).to(torch.bfloat16).float()
        l.lora_A.weight.data = l.lora_A.weight.data.to(torch.bfloat16).float()
    with torch.no_grad():
        ref = blk(inp["visual_x"], inp["visual_context"], inp["visual_t_mod"], inp["visual_freqs"])
        plain = O.dit_block(Pv, "blocks.0", inp["visual_x"], inp["visual_context"], inp["visual_t_mod"],
                            inp["visual_freqs"], cfg["visual_heads"], cfg["eps"])
    assert (ref - plain).abs().max() > 0.05  # the adapters do change the output
    blk.to(torch.bfloat16)
    mine = B.DiTBlock.from_reference(blk)
    assert isinstance(mine.self_attn.q, torch.nn.Linear) and isinstance(mine.cross_attn.o, torch.nn.Linear)
    d = to_dev(inp, device="cpu")
    got = mine(d["visual_x"], d["visual_context"], d["visual_t_mod"], d["visual_freqs"])
    assert_close(got, ref, "LoRA-merged block vs reference LoRA forward", ratio=2e-2, fro=8e-3)
    with pytest.raises(TypeError):
        B.modules.merged_linear(torch.nn.Identity())


def test_guided_update_host_logic(emu):
    from dualforce_b200 import step

    g = torch.Generator().manual_seed(4)
    posi = torch.randn(1, 16, 3, 8, 10, generator=g).to(torch.bfloat16)
    nega = torch.randn(1, 16, 3, 8, 10, generator=g).to(torch.bfloat16)
    lat = torch.randn(1, 16, 3, 8, 10, generator=g)
    ref = O.guided_update(posi, nega, lat, 5.0, 0.9, 0.85)
    got = step.guided_update(posi, nega, lat, 5.0, 0.9, 0.85)
    assert got.dtype == torch.float32 and (got - ref).abs().max() <= 1e-6
    # cfg_scale == 1 branch and the in-place form the denoising loop uses (a persistent [latents | condition] buffer)
    buf = torch.cat([lat, torch.zeros(1, 20, 3, 8, 10)], dim=1)
    out = step.guided_update(posi, None, buf[:, :16], 1.0, 0.9, 0.0, out=buf[:, :16])
    assert out.data_ptr() == buf.data_ptr() and (buf[:, :16] - O.guided_update(posi, None, lat, 1.0, 0.9, 0.0)).abs().max() <= 1e-6
    with pytest.raises(TypeError):
        step.guided_update(posi, nega, lat.to(torch.bfloat16), 5.0, 0.9, 0.85)


def test_end_of_schedule_latents_host_logic(emu, step_case):
    """north_star's second tolerance: end-of-schedule latent cosine similarity.  4 scheduler iterations x 2 CFG
    forwards of the tiny dual tower through the product's denoising loop (persistent model-input buffer, memoised
    prompt work, fused CFG + Euler update) against the oracle's restatement of MOVA.__call__'s loop."""
    from dualforce_b200 import step
    from util import metrics

    cfg, Pv, Pa, Pb, inp, gold, _ = step_case
    vis, aud, bridge, pipe = build_step_towers(cfg, Pv, Pa, Pb, device="cpu")
    g = torch.Generator().manual_seed(12)
    f, h, w = cfg["grid_size"]
    latents = torch.randn(1, 16, f, 2 * h, 2 * w, generator=g)
    condition = torch.randn(1, 20, f, 2 * h, 2 * w, generator=g)
    audio = torch.randn(1, cfg["audio_in_dim"], cfg["audio_len"], generator=g)
    pos = inp["context"].to(torch.bfloat16)
    neg = (torch.randn(pos.shape, generator=g) * 0.5).to(torch.bfloat16)
    sched = O.PairScheduler(num_inference_steps=4)
    assert torch.all(sched.sigmas[:-1] > sched.sigmas[1:]) and abs(float(sched.sigmas[0]) - 1.0) < 1e-6
    ref_v, ref_a = O.denoising_loop(Pv, Pa, Pb, cfg, latents, condition, audio, pos.float(), neg.float(), sched, 5.0)
    got_v, got_a = step.denoising_loop(pipe, latents, condition, audio, pos, neg, sched.get_pairs(),
                                       sched.timestep_to_sigma, cfg["video_fps"], cfg_scale=5.0)
    mv, ma = metrics(got_v, ref_v), metrics(got_a, ref_a)
    assert mv["finite"] and ma["finite"]
    assert mv["cos"] >= 0.999 and ma["cos"] >= 0.999, (mv, ma)
    assert mv["rel_fro"] <= 3e-2 and ma["rel_fro"] <= 3e-2, (mv, ma)
    assert (ref_v - latents).abs().max() > 0.5  # the schedule moved the latents
    # the caller's tensors are untouched; prompt work ran once per prompt, not once per step
    assert emu["patchify"] == 4 * 2 * 2 and emu["cfg_euler_step"] == 4 * 2
    n_blocks = cfg["visual_layers"] + cfg["audio_layers"]
    per_forward = 17 * n_blocks + 14 * min(cfg["visual_layers"], cfg["audio_layers"])
    assert emu["linear"] + emu["layernorm"] + emu["rmsnorm_rope_"] + emu["attention"] + emu["add_to_f32"] < 8 * (per_forward + 12)


def test_cfg_merged_step_equals_two_calls(emu, step_case):
    """SURVEY 8(f)2 / the reference's `cfg_merge` branch (pipeline_mova.py:443-445): [positive, negative] prompts as ONE
    B = 2 forward.  No arithmetic couples the samples, so each half must equal its own B = 1 call -- exactly in this
    emulation -- whether the latents are given once or per sample; and the merged denoising loop must end on the same
    latents as the two-call loop."""
    from dualforce_b200 import step

    cfg, Pv, Pa, Pb, inp, gold, _ = step_case
    vis, aud, bridge, pipe = build_step_towers(cfg, Pv, Pa, Pb, device="cpu")
    g = torch.Generator().manual_seed(5)
    pos = inp["context"].to(torch.bfloat16)
    neg = (torch.randn(pos.shape, generator=g) * 0.5).to(torch.bfloat16)
    kw = dict(visual_dit=vis, visual_latents=inp["visual_latents"], audio_latents=inp["audio_latents"],
              timestep=inp["timestep"], audio_timestep=None, video_fps=cfg["video_fps"])
    pv, pa = pipe.inference_single_step(context=pos, **kw)
    nv, na = pipe.inference_single_step(context=neg, **kw)
    both = torch.cat([pos, neg], dim=0)
    bv, ba = pipe.inference_single_step(context=both, **kw)
    emu.clear()
    pipe.inference_single_step(context=both, **kw)  # prompt memos warm, like the single call counted below
    merged_calls = dict(emu)
    assert bv.shape == (2,) + tuple(pv.shape[1:]) and ba.shape == (2,) + tuple(pa.shape[1:])
    assert torch.equal(bv[0:1], pv) and torch.equal(bv[1:2], nv) and torch.equal(ba[0:1], pa) and torch.equal(ba[1:2], na)
    # per-sample latents (what the reference's chunk(2) contract implies) give the same result
    bv2, ba2 = pipe.inference_single_step(context=torch.cat([pos, neg], dim=0),
                                          **dict(kw, visual_latents=inp["visual_latents"].repeat(2, 1, 1, 1, 1),
                                                 audio_latents=inp["audio_latents"].repeat(2, 1, 1)))
    assert torch.equal(bv2, bv) and torch.equal(ba2, ba)
    # the batch really is batched: one GEMM / LayerNorm / attention launch per site, not one per sample
    emu.clear()
    pipe.inference_single_step(context=pos, **kw)
    single_calls = dict(emu)
    for name in ("linear", "layernorm", "attention"):
        assert merged_calls[name] == single_calls[name], (name, merged_calls[name], single_calls[name])
    # the loop of MOVA.__call__ with cfg_merge: same final latents as two calls per iteration
    f, h, w = cfg["grid_size"]
    latents = torch.randn(1, 16, f, 2 * h, 2 * w, generator=g)
    condition = torch.randn(1, 20, f, 2 * h, 2 * w, generator=g)
    audio = torch.randn(1, cfg["audio_in_dim"], cfg["audio_len"], generator=g)
    sched = O.PairScheduler(num_inference_steps=3)
    args = (pipe, latents, condition, audio, pos, neg, sched.get_pairs(), sched.timestep_to_sigma, cfg["video_fps"])
    two_v, two_a = step.denoising_loop(*args, cfg_scale=5.0)
    one_v, one_a = step.denoising_loop(*args, cfg_scale=5.0, cfg_merge=True)
    assert torch.equal(two_v, one_v) and torch.equal(two_a, one_a)
    # context parallelism keeps the two-call form
    with pytest.raises(NotImplementedError):
        pipe.forward_dual_tower_dit(vis, torch.zeros(2, 4, 256, dtype=torch.bfloat16), None, None, None, None, None,
                                    torch.zeros(4, 64, dtype=torch.complex64), None, (1, 2, 2), 24.0, cp_mesh=object())


def test_split_kv_bridge_attention_host_logic(emu):
    """Few queries against many keys: the keys are cut in equal chunks run as the batch dimension of one launch, then
    merged exactly with their log-sum-exps (modules._kv_splits picks the factor from the shape alone)."""
    import dualforce_b200 as B
    from dualforce_b200.modules import _kv_splits

    assert _kv_splits(403, 43120, 12) > 1 and 43120 % _kv_splits(403, 43120, 12) == 0  # the v2a bridge shape
    assert _kv_splits(43120, 403, 40) == 1 and _kv_splits(403, 4400, 12) == 1 and _kv_splits(403, 8209, 12) == 1
    g = torch.Generator().manual_seed(2)
    dim, kv_dim, H = 256, 384, 2
    cca = B.ConditionalCrossAttention(dim, kv_dim, H).to(torch.bfloat16)
    for prm in cca.parameters():
        prm.data = (torch.randn(prm.shape, generator=g) * (0.05 if prm.dim() > 1 else 0.1)).to(torch.bfloat16)
    x = torch.randn(1, 7, dim, generator=g).to(torch.bfloat16)
    y = torch.randn(1, 8192, kv_dim, generator=g).to(torch.bfloat16)
    n = _kv_splits(7, 8192, H)
    assert n > 1
    split = cca(x, y)
    assert emu["lse_merge"] == 1 and emu["attention"] == 1
    q = cca.project_q(x, None)
    k, v = cca.project_kv(y, None)
    plain = B.ops.linear(B.ops.attention(q, k, v, H), cca.o.weight, cca.o.bias)
    assert emu["lse_merge"] == 1 and emu["attention"] == 2
    assert_close(split, plain.float(), "split-KV vs unsplit", ratio=1e-2, fro=6e-3)
