"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs, against the committed
golden outputs of the reference, and -- at BASELINE.json sizes -- through size-independent properties.

Tolerances (bf16 path vs fp32 oracle fed the same bf16-rounded weights and inputs) are in tests/util.py."""
import json
import math
import os

import numpy as np
import pytest
import torch

import mova_oracle as O
from util import (TOL_DELTA_COS, assert_close, bf16_round, build_towers, metrics, to_dev)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ops():
    import dualforce_b200 as B

    B._lib.require_device(0)
    return B.ops


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("M,N,K", [(1, 8, 8), (77, 136, 72), (403, 1536, 1536), (1000, 5120, 1536), (4400, 3072, 5120)])
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_linear(ops, M, N, K, epi):
    x, w, b = rnd(M, K, seed=1), rnd(N, K, scale=1 / math.sqrt(K), seed=2), rnd(N, scale=0.1, seed=3)
    res, gate = rnd(M, N, seed=4), torch.randn(N, generator=torch.Generator().manual_seed(5))
    ref = x.float() @ w.float().t() + b.float()
    if epi == 1:
        ref = O.gelu_tanh(ref)
    if epi == 2:
        ref = res.float() + gate * 0.7 * ref
    got = ops.linear(x.cuda(), w.cuda(), b.cuda(), epilogue=epi, residual=res.cuda() if epi == 2 else None,
                     gate=gate.cuda() if epi == 2 else None, scale=0.7 if epi == 2 else 1.0)
    assert_close(got, ref, f"linear {M}x{N}x{K} epi{epi}", ratio=6e-3, fro=4e-3)


def test_linear_strided_and_inplace_residual(ops):
    buf = rnd(300, 3 * 256, seed=1).cuda()
    x = buf[:, 256:512]  # column slice: row stride 768
    w, b = rnd(256, 256, scale=1 / 16, seed=2).cuda(), rnd(256, scale=0.1, seed=3).cuda()
    r = rnd(300, 256, seed=4).cuda()
    ref = r.float().cpu() + (x.float().cpu() @ w.float().cpu().t() + b.float().cpu())
    out = ops.linear(x, w, b, epilogue=2, residual=r, out=r)
    assert out.data_ptr() == r.data_ptr()
    assert_close(out, ref, "in-place residual", ratio=6e-3, fro=4e-3)


@pytest.mark.parametrize("segs", [2, 8])
def test_linear_segmented_operands(ops, segs):
    M, K, N = 333, 1536, 3 * 512 * 2
    x = rnd(M, K, seed=1)
    w, b = rnd(N, K, scale=1 / math.sqrt(K), seed=2), rnd(N, scale=0.1, seed=3)
    ref = x.float() @ w.float().t() + b.float()
    # A given source-rank-major [segs, M, K/segs]
    xs = x.reshape(M, segs, K // segs).permute(1, 0, 2).contiguous().cuda()
    got = ops.linear(xs, w.cuda(), b.cuda(), segments=segs)
    assert_close(got, ref, "segmented A", ratio=6e-3, fro=4e-3)
    # C written destination-rank-major [segs, M, N/segs]
    got2 = ops.linear(x.cuda(), w.cuda(), b.cuda(), out_segments=segs)
    assert got2.shape == (segs, M, N // segs)
    assert_close(got2.permute(1, 0, 2).reshape(M, N), ref, "segmented C", ratio=6e-3, fro=4e-3)


@pytest.mark.parametrize("L,d", [(1, 128), (77, 5120), (403, 1536)])
@pytest.mark.parametrize("affine,mod", [(False, False), (True, False), (False, True), (True, True)])
def test_layernorm(ops, L, d, affine, mod):
    x = rnd(L, d, scale=2.0, seed=1)
    w, b = rnd(d, seed=2), rnd(d, seed=3)
    g = torch.Generator().manual_seed(4)
    shift, scale = torch.randn(d, generator=g), torch.randn(d, generator=g) * 0.5
    ref = O.layer_norm(x.float(), 1e-6, w.float() if affine else None, b.float() if affine else None)
    if mod:
        ref = O.modulate(ref, shift, scale)
    got = ops.layernorm(x.cuda(), 1e-6, weight=w.cuda() if affine else None, bias=b.cuda() if affine else None,
                        shift=shift.cuda() if mod else None, scale=scale.cuda() if mod else None)
    assert_close(got, ref, "layernorm", ratio=5e-3, fro=4e-3)


@pytest.mark.parametrize("L,H", [(60, 2), (403, 12), (50, 40)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_rmsnorm_rope(ops, L, H, mode):
    d = H * 128
    x, w = rnd(1, L, d, scale=2.0, seed=1), (1 + 0.1 * rnd(d, seed=2).float()).to(torch.bfloat16)
    ref = O.rms_norm(x.float(), w.float(), 1e-6)
    cos = sin = None
    if mode == 1:
        fr = O.video_freqs(128, (L, 1, 1))
        ref = O.rope_interleaved(ref, fr, 128)
        cos, sin = fr.real.float().reshape(L, 64).cuda().contiguous(), fr.imag.float().reshape(L, 64).cuda().contiguous()
    elif mode == 2:
        (c, s), _ = O.build_aligned_freqs(24.0, (L, 1, 1), 4, 50.0)
        ref = O.rope_half(ref, c, s, 128)
        cos, sin = c[0].cuda().contiguous(), s[0].cuda().contiguous()
    xg = x[0].cuda().clone()
    ops.rmsnorm_rope_(xg, w.cuda(), 1e-6, cos=cos, sin=sin, rope_mode=mode)
    assert_close(xg, ref[0], "rmsnorm_rope", ratio=6e-3, fro=4e-3)


def test_rmsnorm_rope_segmented(ops):
    L, H, segs = 37, 8, 4
    d, wseg = H * 128, H * 128 // segs
    x, w = rnd(1, L, d, scale=2.0, seed=1), (1 + 0.1 * rnd(d, seed=2).float()).to(torch.bfloat16)
    fr = O.video_freqs(128, (L, 1, 1))
    ref = O.rope_interleaved(O.rms_norm(x.float(), w.float(), 1e-6), fr, 128)[0]
    # segmented storage: [segs, L, 3*wseg], q part = columns [0, wseg)
    buf = torch.zeros(segs, L, 3 * wseg, dtype=torch.bfloat16)
    buf[:, :, :wseg] = x[0].reshape(L, segs, wseg).permute(1, 0, 2)
    buf = buf.cuda()
    ops.rmsnorm_rope_(buf[0][:, :wseg], w.cuda(), 1e-6, cos=fr.real.float().reshape(L, 64).cuda().contiguous(),
                      sin=fr.imag.float().reshape(L, 64).cuda().contiguous(), rope_mode=1, segments=segs,
                      seg_stride=L * 3 * wseg)
    got = buf[:, :, :wseg].permute(1, 0, 2).reshape(L, d)
    assert_close(got, ref, "segmented rmsnorm_rope", ratio=6e-3, fro=4e-3)
    assert float(buf[:, :, wseg:].abs().max()) == 0.0  # neighbours untouched


@pytest.mark.parametrize("B,Sq,Skv,H", [(1, 1, 1, 1), (1, 36, 36, 12), (2, 300, 403, 3), (1, 403, 403, 12),
                                         (1, 1000, 512, 4), (1, 403, 4400, 12), (1, 129, 129, 2), (1, 2000, 2000, 5)])
def test_attention(ops, B, Sq, Skv, H):
    q, k, v = rnd(B, Sq, H * 128, seed=1, scale=1.5), rnd(B, Skv, H * 128, seed=2, scale=1.5), rnd(B, Skv, H * 128, seed=3)
    ref, lse_ref = O.attention(q.float(), k.float(), v.float(), H, return_lse=True)
    got, lse = ops.attention(q.cuda(), k.cuda(), v.cuda(), H, return_lse=True)
    assert_close(got, ref, "attention", ratio=1e-2, fro=6e-3)
    assert (lse.cpu() - lse_ref).abs().max() < 2e-3


def test_attention_on_fused_qkv_views(ops):
    S, H = 520, 3
    d = H * 128
    qkv = rnd(1, S, 3 * d, seed=1, scale=1.5)
    ref = O.attention(qkv[..., :d].float(), qkv[..., d:2 * d].float(), qkv[..., 2 * d:].float(), H)
    g = qkv.cuda()
    got = ops.attention(g[..., :d], g[..., d:2 * d], g[..., 2 * d:], H)
    assert_close(got, ref, "attention on views", ratio=1e-2, fro=6e-3)


def test_attention_peaked_softmax_rescales(ops):
    """Scores that grow along the key axis force the lazy-rescale branch (running max jumps by > 2^8)."""
    S, H = 1024, 1
    q = torch.zeros(1, S, 128)
    q[..., 0] = 8.0
    k = torch.zeros(1, S, 128)
    k[0, :, 0] = torch.linspace(0, 40, S)  # score = 8*k0/sqrt(128): up to ~28 nats, spread over 8 key blocks
    v = rnd(1, S, 128, seed=3).float()
    ref = O.attention(q, k, v, H)
    got = ops.attention(q.bfloat16().cuda(), k.bfloat16().cuda(), v.bfloat16().cuda(), H)
    ref = O.attention(q.bfloat16().float(), k.bfloat16().float(), v.bfloat16().float(), H)
    assert_close(got, ref, "peaked attention", ratio=1e-2, fro=8e-3)


def test_lse_merge_equals_full_attention(ops):
    S, Skv, H = 403, 4400, 12
    q, k, v = rnd(1, S, H * 128, seed=1, scale=1.5).cuda(), rnd(1, Skv, H * 128, seed=2, scale=1.5).cuda(), rnd(1, Skv, H * 128, seed=3).cuda()
    full = ops.attention(q, k, v, H)
    parts = [ops.attention(q, k[:, a:b], v[:, a:b], H, return_lse=True) for a, b in ((0, 1100), (1100, 1101), (1101, 4400))]
    merged = ops.lse_merge(torch.stack([o[0] for o, _ in parts]), torch.stack([l[0] for _, l in parts]), H)
    assert_close(merged, full[0], "lse merge", ratio=1e-2, fro=5e-3)


# ------------------------------------------------------------------------------------------------ modules
def _case(cfg, seed):
    Pv, Pa, Pb, inp = O.make_case(cfg, seed)
    return bf16_round(Pv), bf16_round(Pa), bf16_round(Pb), bf16_round(inp)


def _check_block(got, ref, x, name):
    m = assert_close(got, ref, name)
    dcos = metrics(got.float().cpu() - x, ref - x)["cos"]
    assert dcos >= TOL_DELTA_COS, f"{name}: residual-delta cosine {dcos}"
    return m


def test_tiny_modules_vs_oracle():
    cfg = O.TINY_CFG
    Pv, Pa, Pb, inp = _case(cfg, 1234)
    vis, aud, bridge, pipe = build_towers(cfg, Pv, Pa, Pb)
    d = to_dev(inp)
    y = vis.blocks[0](d["visual_x"], d["visual_context"], d["visual_t_mod"], d["visual_freqs"])
    ref = O.dit_block(Pv, "blocks.0", inp["visual_x"], inp["visual_context"], inp["visual_t_mod"], inp["visual_freqs"],
                      cfg["visual_heads"], cfg["eps"])
    _check_block(y, ref, inp["visual_x"], "tiny video block")
    y = aud.blocks[0](d["audio_x"], d["audio_context"], d["audio_t_mod"], d["audio_freqs"])
    ref = O.dit_block(Pa, "blocks.0", inp["audio_x"], inp["audio_context"], inp["audio_t_mod"], inp["audio_freqs"],
                      cfg["audio_heads"], cfg["eps"])
    _check_block(y, ref, inp["audio_x"], "tiny audio block")
    v_cs, a_cs = O.build_aligned_freqs(cfg["video_fps"], cfg["grid_size"], cfg["audio_len"], cfg["audio_fps"])
    gv, ga = bridge.build_aligned_freqs(cfg["video_fps"], cfg["grid_size"], cfg["audio_len"], device=torch.device("cuda"),
                                        dtype=torch.float32)
    assert (gv[0].cpu() - v_cs[0]).abs().max() < 2e-5 and (ga[1].cpu() - a_cs[1]).abs().max() < 2e-5
    bv, ba = bridge(0, d["visual_x"], d["audio_x"], x_freqs=gv, y_freqs=ga, condition_scale=1.0,
                    video_grid_size=cfg["grid_size"])
    rv, ra = O.bridge_layer(Pb, 0, inp["visual_x"], inp["audio_x"], v_cs, a_cs, cfg["head_dim"])
    _check_block(bv, rv, inp["visual_x"], "tiny bridge a2v")
    _check_block(ba, ra, inp["audio_x"], "tiny bridge v2a")
    # inputs are not mutated (the same context / t_mod / hidden states are reused across layers and directions)
    assert torch.equal(d["visual_x"].cpu(), inp["visual_x"].bfloat16())


def test_forward_dual_tower_vs_golden_and_oracle():
    with open(os.path.join(GOLDEN, "tiny_dual_tower.json")) as f:
        meta = json.load(f)
    cfg = dict(meta["cfg"], grid_size=tuple(meta["cfg"]["grid_size"]))
    gold = np.load(os.path.join(GOLDEN, "tiny_dual_tower.npz"))
    Pv, Pa, Pb, inp = _case(cfg, meta["seed"])
    vis, aud, bridge, pipe = build_towers(cfg, Pv, Pa, Pb)
    d = to_dev(inp)
    fv, fa = pipe.forward_dual_tower_dit(vis, d["visual_x"], d["audio_x"], d["visual_context"], d["audio_context"],
                                         d["visual_t_mod"], d["audio_t_mod"], d["visual_freqs"], d["audio_freqs"],
                                         cfg["grid_size"], cfg["video_fps"])
    rv, ra = O.forward_dual_tower_dit(Pv, Pa, Pb, cfg, inp["visual_x"], inp["audio_x"], inp["visual_context"],
                                      inp["audio_context"], inp["visual_t_mod"], inp["audio_t_mod"], inp["visual_freqs"],
                                      inp["audio_freqs"], cfg["grid_size"], cfg["video_fps"])
    # vs the oracle on identical (bf16-rounded) weights: compute error only, accumulated over 3+2 layers
    assert_close(fv, rv, "dual tower visual vs oracle", ratio=3e-2, fro=1.2e-2)
    assert_close(fa, ra, "dual tower audio vs oracle", ratio=3e-2, fro=1.2e-2)
    # vs the reference's own fp32 outputs (fp32 weights): adds the bf16 rounding of weights and inputs
    assert_close(fv, torch.from_numpy(gold["final_visual"]), "dual tower visual vs golden", ratio=4e-2, fro=2e-2)
    assert_close(fa, torch.from_numpy(gold["final_audio"]), "dual tower audio vs golden", ratio=4e-2, fro=2e-2)
    assert metrics(fv, torch.from_numpy(gold["final_visual"]))["cos"] > 0.999


@pytest.mark.parametrize("dim,heads,ffn,L", [(5120, 40, 13824, 640), (1536, 12, 8960, 403), (1536, 12, 8960, 36)])
def test_full_width_block_vs_oracle(dim, heads, ffn, L):
    """MOVA widths (video 5120/40/13824, audio 1536/12/8960) at a token count the CPU oracle finishes in seconds."""
    gen = torch.Generator().manual_seed(5)
    P = bf16_round(O.make_block_weights(gen, "blocks.0", dim, ffn))
    cfg = dict(O.TINY_CFG, visual_dim=dim, visual_heads=heads, visual_ffn=ffn, visual_layers=1, audio_layers=0)
    x = rnd(1, L, dim, seed=6).float()
    ctx = rnd(1, 512, dim, seed=7).float()
    ctx[:, 64:] = 0
    t_mod = rnd(1, 6, dim, seed=8, scale=0.3).float()
    freqs = O.video_freqs(128, (L // 4, 2, 2)) if L % 4 == 0 else O.audio_freqs(128, L)
    import dualforce_b200 as B

    blk = B.DiTBlock(False, dim, heads, ffn, 1e-6)
    blk.load_state_dict({k[len("blocks.0."):]: v for k, v in P.items()})
    blk.to("cuda", torch.bfloat16)
    got = blk(x.bfloat16().cuda(), ctx.bfloat16().cuda(), t_mod.bfloat16().cuda(), freqs.cuda())
    ref = O.dit_block(P, "blocks.0", x, ctx, t_mod, freqs, heads)
    _check_block(got, ref, x, f"block {dim}/{heads} L={L}")
    del cfg


def test_full_width_bridge_vs_oracle():
    cfg = dict(O.REDUCED_360P_CFG, grid_size=(3, 8, 10), audio_len=36)
    gen = torch.Generator().manual_seed(11)
    Pb = bf16_round(O.make_bridge_weights(gen, [0], 5120, 1536))
    xv, xa = rnd(1, 240, 5120, seed=1).float(), rnd(1, 36, 1536, seed=2).float()
    import dualforce_b200 as B

    bridge = B.DualTowerConditionalBridge(visual_layers=1, audio_layers=1, visual_hidden_dim=5120, audio_hidden_dim=1536,
                                          audio_fps=50.0, head_dim=128, interaction_strategy="full", apply_cross_rope=True)
    bridge.load_state_dict(Pb)
    bridge.to("cuda", torch.bfloat16)
    v_cs, a_cs = O.build_aligned_freqs(24.0, cfg["grid_size"], 36, 50.0)
    gv, ga = bridge.build_aligned_freqs(24.0, cfg["grid_size"], 36, device=torch.device("cuda"), dtype=torch.float32)
    bv, ba = bridge(0, xv.bfloat16().cuda(), xa.bfloat16().cuda(), x_freqs=gv, y_freqs=ga, condition_scale=1.0)
    rv, ra = O.bridge_layer(Pb, 0, xv, xa, v_cs, a_cs)
    _check_block(bv, rv, xv, "bridge a2v 5120<-1536")
    _check_block(ba, ra, xa, "bridge v2a 1536<-5120")


class _LoRAWrapped(torch.nn.Module):
    """Stand-in with the attribute names of the reference's LoRALinear (engine/trainer/accelerate/lora_utils.py:19-90:
    ``original_layer``, ``lora_A``, ``lora_B``, ``scaling``) -- what ``modules.merged_linear`` duck-types on; the real
    class is exercised on CPU in tests/test_host_emulated.py (there is no reference tree on the GPU box)."""

    def __init__(self, original_layer, rank, alpha, gen):
        super().__init__()
        self.original_layer = original_layer
        self.scaling = alpha / rank
        self.lora_A = torch.nn.Linear(original_layer.in_features, rank, bias=False)
        self.lora_B = torch.nn.Linear(rank, original_layer.out_features, bias=False)
        self.lora_A.weight.data = (torch.randn(self.lora_A.weight.shape, generator=gen) * 0.1).to(torch.bfloat16).float()
        self.lora_B.weight.data = (torch.randn(self.lora_B.weight.shape, generator=gen) * 0.2).to(torch.bfloat16).float()


def test_lora_adapters_are_folded_at_install_on_device():
    """SURVEY 8f.4 on the device: a block whose q / k / v / o carry LoRA adapters (h = W x + (alpha / r) B A x,
    mova_lora.py:147-188) is swapped with the adapters folded into the weights; the CUDA block matches the oracle
    evaluated with the un-merged formula."""
    import dualforce_b200 as B

    cfg = O.TINY_CFG
    Pv, Pa, Pb, inp = O.make_case(cfg, 5)
    Pv, inp = bf16_round(Pv), bf16_round(inp)
    host = B.DiTBlock(False, cfg["visual_dim"], cfg["visual_heads"], cfg["visual_ffn"], cfg["eps"])
    host.load_state_dict({k[len("blocks.0."):]: v for k, v in Pv.items() if k.startswith("blocks.0.")})
    gen = torch.Generator().manual_seed(9)
    P_lora = dict(Pv)
    for attn in ("self_attn", "cross_attn"):
        mod = getattr(host, attn)
        for name in ("q", "k", "v", "o"):
            wrapped = _LoRAWrapped(getattr(mod, name), rank=4, alpha=8.0, gen=gen)
            setattr(mod, name, wrapped)
            key = f"blocks.0.{attn}.{name}.weight"
            P_lora[key] = Pv[key] + wrapped.scaling * (wrapped.lora_B.weight.data @ wrapped.lora_A.weight.data)
    host.to(torch.bfloat16)
    mine = B.DiTBlock.from_reference(host).to("cuda")
    assert isinstance(mine.self_attn.q, torch.nn.Linear) and isinstance(mine.cross_attn.o, torch.nn.Linear)
    d = to_dev(inp)
    got = mine(d["visual_x"], d["visual_context"], d["visual_t_mod"], d["visual_freqs"])
    ref = O.dit_block(P_lora, "blocks.0", inp["visual_x"], inp["visual_context"], inp["visual_t_mod"], inp["visual_freqs"],
                      cfg["visual_heads"], cfg["eps"])
    plain = O.dit_block(Pv, "blocks.0", inp["visual_x"], inp["visual_context"], inp["visual_t_mod"], inp["visual_freqs"],
                        cfg["visual_heads"], cfg["eps"])
    assert (ref - plain).abs().max() > 0.05  # the adapters do change the output
    assert_close(got, ref, "LoRA-merged block on the device vs un-merged oracle", ratio=2e-2, fro=8e-3)


def test_bridge_rope_reference_bf16_mode_vs_reference_bf16_run():
    """``bridge_rope="reference_bf16"`` on the device against the reference's own bf16 run of the bridge
    (tests/golden/bridge_rope_bf16.npz: reference modules after ``.to(torch.bfloat16)``, 403 audio positions): the
    tables are bit-identical to the reference's, the layer output tracks the reference's bf16 output, and the exact-
    frequency mode is measurably further from it (the reference's rounded ``inv_freq`` moves the phases by up to
    0.8 rad) -- while still matching the fp32 reference run."""
    import dualforce_b200 as B
    from test_oracle_golden import load_bridge_bf16_case

    cfg, Pv, Pa, Pb, inp, gold = load_bridge_bf16_case()
    bridge = B.DualTowerConditionalBridge(
        visual_layers=cfg["visual_layers"], audio_layers=cfg["audio_layers"], visual_hidden_dim=cfg["visual_dim"],
        audio_hidden_dim=cfg["audio_dim"], audio_fps=cfg["audio_fps"], head_dim=cfg["head_dim"],
        interaction_strategy=cfg["interaction_strategy"], apply_cross_rope=True)
    bridge.load_state_dict(Pb)
    bridge.to("cuda", torch.bfloat16)
    xv, xa = inp["visual_x"].to(torch.bfloat16), inp["audio_x"].to(torch.bfloat16)
    kw = dict(video_fps=cfg["video_fps"], grid_size=cfg["grid_size"], audio_steps=cfg["audio_len"],
              device=torch.device("cuda"), dtype=torch.float32)
    errs = {}
    for mode in ("reference_bf16", "fp32"):
        bridge.bridge_rope = mode
        gv, ga = bridge.build_aligned_freqs(**kw)
        if mode == "reference_bf16":
            assert torch.equal(ga[0].cpu(), gold["cos_a"]) and torch.equal(gv[1].cpu(), gold["sin_v"])
        bv, ba = bridge(0, xv.cuda(), xa.cuda(), x_freqs=gv, y_freqs=ga, condition_scale=1.0)
        for name, got, x in (("visual", bv, xv), ("audio", ba, xa)):
            d = got.float().cpu() - x.float()
            g16 = gold[f"bridge0_{name}"] - x.float()
            g32 = gold[f"bridge0_{name}_fp32"] - x.float()
            errs[(mode, name)] = (((d - g16).norm() / g16.norm()).item(), ((d - g32).norm() / g32.norm()).item())
    # vs the reference's bf16 run: bf16 arithmetic noise on both sides in the matching mode, RoPE mismatch otherwise
    assert errs[("reference_bf16", "audio")][0] < 0.04 and errs[("reference_bf16", "visual")][0] < 0.04, errs
    assert errs[("fp32", "audio")][0] > 2 * errs[("reference_bf16", "audio")][0], errs
    # vs the reference's fp32 run it is the other way round
    assert errs[("fp32", "audio")][1] < 0.03 and errs[("fp32", "audio")][1] < errs[("reference_bf16", "audio")][1], errs


# ------------------------------------------------------------------------------------------------ full-size properties
def test_attention_properties_at_360p_size(ops):
    """BASELINE.json configs[1] geometry (L_v = 43120, head_dim 128), 2 of the 40 heads to bound memory/time:
    (1) V = 1 gives exactly 1 (softmax rows sum to one), (2) attention is linear in V,
    (3) split-KV + LSE merge reproduces the unsplit result."""
    S, H = 43120, 2
    q, k = rnd(1, S, H * 128, seed=1, scale=1.2).cuda(), rnd(1, S, H * 128, seed=2, scale=1.2).cuda()
    ones = torch.ones(1, S, H * 128, dtype=torch.bfloat16, device="cuda")
    o1 = ops.attention(q, k, ones, H)
    assert (o1.float() - 1.0).abs().max() < 8e-3
    v1, v2 = rnd(1, S, H * 128, seed=3).cuda(), rnd(1, S, H * 128, seed=4).cuda()
    a, b, ab = ops.attention(q, k, v1, H), ops.attention(q, k, v2, H), ops.attention(q, k, (v1.float() + v2.float()).bfloat16(), H)
    m = metrics(ab, a.float() + b.float())
    assert m["rel_fro"] < 2e-2, m  # bf16 rounding of the two outputs and of v1+v2
    half = S // 2
    parts = [ops.attention(q, k[:, s:e], v1[:, s:e], H, return_lse=True) for s, e in ((0, half), (half, S))]
    merged = ops.lse_merge(torch.stack([o[0] for o, _ in parts]), torch.stack([l[0] for _, l in parts]), H)
    m = metrics(merged, a[0])
    assert m["rel_fro"] < 1e-2, m


def test_cuda_graph_replay_matches_eager():
    """install(..., cuda_graph=True): the captured forward must track changing inputs (static buffers are refreshed,
    RoPE tables are rebuilt inside the graph) and reproduce the eager result bit for bit."""
    cfg = O.TINY_CFG
    Pv, Pa, Pb, inp = _case(cfg, 4321)
    vis, aud, bridge, pipe = build_towers(cfg, Pv, Pa, Pb)

    def run(d):
        return pipe.forward_dual_tower_dit(vis, d["visual_x"], d["audio_x"], d["visual_context"], d["audio_context"],
                                           d["visual_t_mod"], d["audio_t_mod"], d["visual_freqs"], d["audio_freqs"],
                                           cfg["grid_size"], cfg["video_fps"])

    d1 = to_dev(inp)
    d2 = dict(d1)
    d2["visual_x"] = (d1["visual_x"].float() * 0.5 + 0.25).bfloat16()
    d2["visual_freqs"] = torch.roll(d1["visual_freqs"], 3, dims=0).contiguous()  # a different table, same shape
    eager = [run(d1), run(d2)]
    pipe.mova_b200_cuda_graph = True
    graphed = [run(d1), run(d2), run(d1)]
    assert len(pipe._mova_b200_graphs) == 1  # one capture, three replays
    for (ev, ea), (gv, ga) in zip(eager + [eager[0]], graphed):
        assert torch.equal(ev, gv) and torch.equal(ea, ga)
    assert not torch.equal(graphed[0][0], graphed[1][0])
