"""Torch-tensor front end of the C ABI: each function validates layouts, passes raw device pointers and the
current CUDA stream to ``libmova_b200.so`` and returns torch tensors.  No arithmetic happens in Python and there
is no fallback: a CPU tensor, a wrong dtype or a missing library raises.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import EPI_BIAS, EPI_GELU_TANH, EPI_RESIDUAL, ROPE_HALF, ROPE_INTERLEAVED, ROPE_NONE  # noqa: F401

__all__ = [
    "linear", "attention", "layernorm", "rmsnorm_rope_", "lse_merge", "add_to_f32", "patchify", "unpatchify",
    "sinusoidal_embedding", "gemv_f32", "cfg_euler_step",
    "EPI_BIAS", "EPI_GELU_TANH", "EPI_RESIDUAL", "ROPE_NONE", "ROPE_INTERLEAVED", "ROPE_HALF",
]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need(t: torch.Tensor, dtype: torch.dtype, name: str) -> None:
    if not t.is_cuda:
        raise _lib.MovaB200Error(f"{name} must be a CUDA tensor (dualforce_b200 has no CPU path); got {t.device}")
    if t.dtype != dtype:
        raise _lib.MovaB200Error(f"{name} must be {dtype}, got {t.dtype}")
    _lib.require_device(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _rows2d(t: torch.Tensor, name: str) -> Tuple[int, int, int]:
    """(rows, cols, leading dimension) of a tensor viewed as a row-major matrix whose last dim is contiguous."""
    if t.dim() < 2:
        raise _lib.MovaB200Error(f"{name}: need at least 2 dims, got shape {tuple(t.shape)}")
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        raise _lib.MovaB200Error(f"{name}: last dimension must be contiguous (strides {t.stride()})")
    cols = t.shape[-1]
    rows = t.numel() // cols if cols else 0
    ld = t.stride(-2)
    # leading dims must collapse onto one row stride
    expect = ld
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] != 1 and t.stride(d) != expect:
            raise _lib.MovaB200Error(f"{name}: shape {tuple(t.shape)} strides {t.stride()} is not a strided matrix")
        expect *= t.shape[d]
    return rows, cols, ld


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *, epilogue: int = EPI_BIAS,
           residual: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None, scale: float = 1.0,
           out: Optional[torch.Tensor] = None, cta_group: int = 0, segments: int = 1,
           out_segments: int = 1) -> torch.Tensor:
    """``epi(x @ weight.T + bias)`` -- nn.Linear (wan_video_dit.py:171-174) with the GELU-tanh (:270-271) or
    gated-residual (:254-255, interactionv2.py:535) step fused into the tcgen05 GEMM epilogue.

    x: [..., K] bf16 (row stride arbitrary), weight: [N, K] bf16, bias: [N] bf16, gate: [N] fp32,
    residual: [..., N] bf16 (may be ``out``).  Returns [..., N] bf16.

    ``segments > 1``: x is ``[segments, M, K/segments]`` (contiguous) and is read as the matrix
    ``A[m, s*K/segments + c] = x[s, m, c]`` -- the layout the context-parallel all-to-all delivers.
    ``out_segments > 1``: the result is written as ``out[s, m, c] = C[m, s*N/out_segments + c]`` into a contiguous
    ``[out_segments, M, N/out_segments]`` buffer -- the layout the context-parallel all-to-all sends.
    """
    _need(x, torch.bfloat16, "x")
    _need(weight, torch.bfloat16, "weight")
    if segments > 1:
        if x.dim() != 3 or x.shape[0] != segments or not x.is_contiguous():
            raise _lib.MovaB200Error(f"linear: segmented x must be contiguous [segments, M, seg_k], got {tuple(x.shape)}")
        M, seg_k = x.shape[1], x.shape[2]
        K, lda, seg_stride = seg_k * segments, seg_k, M * seg_k
        out_shape = (M,)
    else:
        M, K, lda = _rows2d(x, "x")
        seg_k, seg_stride = K, 0
        out_shape = tuple(x.shape[:-1])
    N, Kw, ldw = _rows2d(weight, "weight")
    if Kw != K:
        raise _lib.MovaB200Error(f"linear: x has K={K}, weight has K={Kw}")
    if out_segments > 1:
        if N % out_segments:
            raise _lib.MovaB200Error(f"linear: N={N} not divisible by out_segments={out_segments}")
        seg_n = N // out_segments
        if out is None:
            out = torch.empty(out_segments, M, seg_n, dtype=torch.bfloat16, device=x.device)
        _need(out, torch.bfloat16, "out")
        if tuple(out.shape) != (out_segments, M, seg_n) or not out.is_contiguous():
            raise _lib.MovaB200Error(f"linear: segmented out must be contiguous {(out_segments, M, seg_n)}")
        ldc, c_seg_stride = seg_n, M * seg_n
    else:
        seg_n, c_seg_stride = N, 0
        if out is None:
            out = torch.empty(*out_shape, N, dtype=torch.bfloat16, device=x.device)
        else:
            _need(out, torch.bfloat16, "out")
        Mo, No, ldc = _rows2d(out, "out")
        if (Mo, No) != (M, N):
            raise _lib.MovaB200Error(f"linear: out is {Mo}x{No}, expected {M}x{N}")
    res_ptr, ldr = None, 0
    if epilogue == EPI_RESIDUAL:
        if residual is None:
            raise _lib.MovaB200Error("linear: EPI_RESIDUAL needs `residual`")
        _need(residual, torch.bfloat16, "residual")
        Mr, Nr, ldr = _rows2d(residual, "residual")
        if (Mr, Nr) != (M, N):
            raise _lib.MovaB200Error(f"linear: residual is {Mr}x{Nr}, expected {M}x{N}")
        res_ptr = residual.data_ptr()
    gate_ptr = None
    if gate is not None:
        _need(gate, torch.float32, "gate")
        if gate.numel() != N or not gate.is_contiguous():
            raise _lib.MovaB200Error("linear: gate must be a contiguous fp32 vector of N elements")
        gate_ptr = gate.data_ptr()
    bias_ptr = None
    if bias is not None:
        _need(bias, torch.bfloat16, "bias")
        if bias.numel() != N or not bias.is_contiguous():
            raise _lib.MovaB200Error("linear: bias must be a contiguous bf16 vector of N elements")
        bias_ptr = bias.data_ptr()
    rc = _lib.load().mova_b200_linear_ex(x.data_ptr(), lda, seg_k, seg_stride, weight.data_ptr(), ldw, bias_ptr,
                                         out.data_ptr(), ldc, seg_n, c_seg_stride, M, N, K, epilogue, res_ptr, ldr,
                                         gate_ptr, float(scale), cta_group, _stream())
    _lib.check(rc, "mova_b200_linear")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int, *, return_lse: bool = False,
              softmax_scale: Optional[float] = None, out: Optional[torch.Tensor] = None,
              variant: Optional[int] = None, emu: int = 4):
    """``softmax(q k^T / sqrt(D)) v`` on the flat ``[B, S, H*D]`` layout of flash_attention()
    (wan_video_dit.py:58-91).  q/k/v may be column slices of a fused projection buffer (any row stride).

    Returns ``[B, Sq, H*D]`` bf16 (and ``lse [B, H, Sq]`` fp32, natural log, when ``return_lse``).
    ``variant`` (measurement only, see ``mova_b200_attn_fwd_variant``): 92 / 91 = round-2 schedule as a CTA pair /
    single CTA, 3 = round-1 schedule; default: the library's shipped schedule.
    """
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _need(t, torch.bfloat16, n)
        if t.dim() != 3 or t.stride(2) != 1:
            raise _lib.MovaB200Error(f"attention: {n} must be [B, S, H*D] with a contiguous last dim")
    B, Sq, HD = q.shape
    Skv = k.shape[1]
    if HD % num_heads:
        raise _lib.MovaB200Error(f"attention: channel dim {HD} not divisible by num_heads {num_heads}")
    D = HD // num_heads
    if k.shape != (B, Skv, HD) or v.shape != (B, Skv, HD):
        raise _lib.MovaB200Error(f"attention: k/v shapes {tuple(k.shape)}/{tuple(v.shape)} do not match q {tuple(q.shape)}")
    if out is None:
        out = torch.empty(B, Sq, HD, dtype=torch.bfloat16, device=q.device)
    else:
        _need(out, torch.bfloat16, "out")
        if out.shape != (B, Sq, HD) or out.stride(2) != 1:
            raise _lib.MovaB200Error("attention: bad `out`")
    lse = torch.empty(B, num_heads, Sq, dtype=torch.float32, device=q.device) if return_lse else None
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(D)
    rec = _lib._TIMERS
    if rec is not None:  # bench.py: per-launch device time of the dominant kernel, on the launching stream
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
    args = (q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(), k.stride(0), k.stride(1), v.data_ptr(), v.stride(0),
            v.stride(1), out.data_ptr(), out.stride(0), out.stride(1), lse.data_ptr() if lse is not None else None, B,
            Sq, Skv, num_heads, D, float(scale))
    if variant is None:
        rc = _lib.load().mova_b200_attn_fwd(*args, _stream())
    else:
        rc = _lib.load().mova_b200_attn_fwd_variant(*args, int(variant), int(emu), None, _stream())
    _lib.check(rc, "mova_b200_attn_fwd")
    if rec is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        rec.append((e0, e1, B, Sq, Skv, num_heads, D))
    return (out, lse) if return_lse else out


def layernorm(x: torch.Tensor, eps: float, *, weight: Optional[torch.Tensor] = None,
              bias: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
              scale: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One-pass ``LayerNorm(x) [*weight + bias] [*(1 + scale) + shift]``: nn.LayerNorm + modulate()
    (wan_video_dit.py:94-96, 267-269, 286, 289; interactionv2.py:322, 349).  shift/scale: fp32 [d]."""
    _need(x, torch.bfloat16, "x")
    L, d, ldx = _rows2d(x, "x")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _need(out, torch.bfloat16, "out")
    Lo, do, ldy = _rows2d(out, "out")
    if (Lo, do) != (L, d):
        raise _lib.MovaB200Error("layernorm: out shape mismatch")
    ptr = {}
    for name, t, dt in (("weight", weight, torch.bfloat16), ("bias", bias, torch.bfloat16),
                        ("shift", shift, torch.float32), ("scale", scale, torch.float32)):
        if t is None:
            ptr[name] = None
        else:
            _need(t, dt, name)
            if t.numel() != d or not t.is_contiguous():
                raise _lib.MovaB200Error(f"layernorm: {name} must be a contiguous vector of {d} elements")
            ptr[name] = t.data_ptr()
    rc = _lib.load().mova_b200_layernorm(x.data_ptr(), ldx, out.data_ptr(), ldy, L, d, float(eps), ptr["weight"],
                                         ptr["bias"], ptr["shift"], ptr["scale"], _stream())
    _lib.check(rc, "mova_b200_layernorm")
    return out


def rmsnorm_rope_(x: torch.Tensor, weight: torch.Tensor, eps: float, *, head_dim: int = 128,
                  cos: Optional[torch.Tensor] = None, sin: Optional[torch.Tensor] = None,
                  rope_mode: int = ROPE_NONE, segments: int = 1, seg_stride: int = 0) -> torch.Tensor:
    """In place ``x = RoPE(RMSNorm_d(x) * weight)`` with the RMS taken over all ``d = H*head_dim`` channels
    (torch.nn.RMSNorm(dim), wan_video_dit.py:175-176) and RoPE in the interleaved (wan_video_dit.py:131-137) or
    rotate-half (interactionv2.py:47-72) convention.  cos/sin: fp32 ``[L, 64]`` / ``[L, 128]`` tables.

    ``segments > 1``: ``x`` is the ``[L, d/segments]`` view of segment 0 and segment s starts ``s * seg_stride``
    elements later (the destination-rank-major q/k/v buffer of the context-parallel path)."""
    _need(x, torch.bfloat16, "x")
    _need(weight, torch.bfloat16, "weight")
    if rope_mode != ROPE_NONE and x.dim() == 3 and x.shape[0] > 1 and cos is not None:
        width = head_dim // 2 if rope_mode == ROPE_INTERLEAVED else head_dim
        if cos.numel() == x.shape[1] * width:
            # a batch that shares one position table (CFG pair: the same tokens twice): one launch per sample, the table
            # row is the token index inside the sample
            for b in range(x.shape[0]):
                rmsnorm_rope_(x[b], weight, eps, head_dim=head_dim, cos=cos, sin=sin, rope_mode=rope_mode,
                              segments=segments, seg_stride=seg_stride)
            return x
    L, seg_len, ldx = _rows2d(x, "x")
    d = seg_len * segments
    if weight.numel() != d or not weight.is_contiguous():
        raise _lib.MovaB200Error("rmsnorm_rope_: weight must be a contiguous bf16 vector of d elements")
    cptr = sptr = None
    if rope_mode != ROPE_NONE:
        width = head_dim // 2 if rope_mode == ROPE_INTERLEAVED else head_dim
        for t, n in ((cos, "cos"), (sin, "sin")):
            if t is None:
                raise _lib.MovaB200Error("rmsnorm_rope_: RoPE tables missing")
            _need(t, torch.float32, n)
            if not t.is_contiguous() or t.numel() != L * width:
                raise _lib.MovaB200Error(f"rmsnorm_rope_: {n} must be contiguous fp32 [{L}, {width}], got {tuple(t.shape)}")
        cptr, sptr = cos.data_ptr(), sin.data_ptr()
    rc = _lib.load().mova_b200_rmsnorm_rope_seg(x.data_ptr(), ldx, seg_len, int(seg_stride), L, d, head_dim,
                                                weight.data_ptr(), float(eps), cptr, sptr, rope_mode, _stream())
    _lib.check(rc, "mova_b200_rmsnorm_rope")
    return x


def lse_merge(o_parts: torch.Tensor, lse_parts: torch.Tensor, num_heads: int, *, return_lse: bool = False):
    """Combine partial attention results over disjoint key sets: ``o_parts [P, rows, H*D]`` bf16 and
    ``lse_parts [P, H, rows]`` fp32 -> ``[rows, H*D]`` (split-KV and the context-parallel v2a bridge)."""
    _need(o_parts, torch.bfloat16, "o_parts")
    _need(lse_parts, torch.float32, "lse_parts")
    if not o_parts.is_contiguous() or not lse_parts.is_contiguous() or o_parts.dim() != 3 or lse_parts.dim() != 3:
        raise _lib.MovaB200Error("lse_merge: o_parts [P, rows, H*D] and lse_parts [P, H, rows] must be contiguous")
    P, rows, HD = o_parts.shape
    if lse_parts.shape != (P, num_heads, rows):
        raise _lib.MovaB200Error(f"lse_merge: lse_parts shape {tuple(lse_parts.shape)} != {(P, num_heads, rows)}")
    out = torch.empty(rows, HD, dtype=torch.bfloat16, device=o_parts.device)
    lse = torch.empty(num_heads, rows, dtype=torch.float32, device=o_parts.device) if return_lse else None
    rc = _lib.load().mova_b200_lse_merge(o_parts.data_ptr(), lse_parts.data_ptr(), P, out.data_ptr(), HD,
                                         lse.data_ptr() if lse is not None else None, rows, num_heads,
                                         HD // num_heads, _stream())
    _lib.check(rc, "mova_b200_lse_merge")
    return (out, lse) if return_lse else out


def add_to_f32(a: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``float(a) + float(b)`` for bf16 inputs: ``modulation + t_mod`` of DiTBlock (wan_video_dit.py:279-280),
    kept in fp32 so shift/scale/gate enter the fused kernels unrounded."""
    _need(a, torch.bfloat16, "a")
    a = a.contiguous()
    bptr = None
    if b is not None:
        _need(b, torch.bfloat16, "b")
        b = b.expand_as(a).contiguous()
        bptr = b.data_ptr()
    out = torch.empty(a.shape, dtype=torch.float32, device=a.device)
    rc = _lib.load().mova_b200_add_to_f32(a.data_ptr(), bptr, out.data_ptr(), a.numel(), _stream())
    _lib.check(rc, "mova_b200_add_to_f32")
    return out


# ----------------------------------------------------------------------------------------------------------------
# the step either side of the dual-tower forward (MOVA.inference_single_step, pipeline_mova.py:500-609)
# ----------------------------------------------------------------------------------------------------------------
def patchify(x: torch.Tensor, patch_size, *, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """im2col of the stride == kernel patch embedding: ``x [C, F, H, W]`` (or ``[C, F]`` for the audio Conv1d),
    fp32 or bf16 -> bf16 ``[L, C*pt*ph*pw]`` with tokens in (f, h, w) order and columns in the order of
    ``conv.weight.view(dim, -1)`` (wan_video_dit.py:399-409; wan_audio_dit.py:180-189).  The bf16 cast of
    pipeline_mova.py:556-557 is fused."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise _lib.MovaB200Error(f"patchify: x must be fp32 or bf16, got {x.dtype}")
    _need(x, x.dtype, "x")
    if x.dim() == 2:
        x = x[:, :, None, None]
    if x.dim() != 4 or not x.is_contiguous():
        raise _lib.MovaB200Error(f"patchify: x must be a contiguous [C, F, H, W] (or [C, F]) tensor, got {tuple(x.shape)}")
    p = tuple(int(v) for v in patch_size) + (1, 1)
    pt, ph, pw = p[0], p[1], p[2]
    C, F, H, W = x.shape
    if F % pt or H % ph or W % pw:
        raise _lib.MovaB200Error(f"patchify: latent {F}x{H}x{W} is not a multiple of the patch {pt}x{ph}x{pw}")
    L, K = (F // pt) * (H // ph) * (W // pw), C * pt * ph * pw
    if out is None:
        out = torch.empty(L, K, dtype=torch.bfloat16, device=x.device)
    _need(out, torch.bfloat16, "out")
    Lo, Ko, ldo = _rows2d(out, "out")
    if (Lo, Ko) != (L, K):
        raise _lib.MovaB200Error(f"patchify: out is {Lo}x{Ko}, expected {L}x{K}")
    rc = _lib.load().mova_b200_patchify(x.data_ptr(), int(x.dtype == torch.float32), C, F, H, W, pt, ph, pw,
                                        out.data_ptr(), ldo, _stream())
    _lib.check(rc, "mova_b200_patchify")
    return out


def unpatchify(x: torch.Tensor, grid_size, patch_size, out_channels: int) -> torch.Tensor:
    """``'(f h w) (x y z c) -> c (f x) (h y) (w z)'`` (wan_video_dit.py:411-416; wan_audio_dit.py:191-195 with
    y = z = 1): ``x [L, pt*ph*pw*C]`` bf16 -> ``[C, F*pt, H*ph, W*pw]`` bf16 (``[C, F*pt]`` for a 1-D grid)."""
    _need(x, torch.bfloat16, "x")
    L, cols, ldi = _rows2d(x, "x")
    g = tuple(int(v) for v in grid_size)
    one_d = len(g) == 1
    g = g + (1, 1)
    p = tuple(int(v) for v in patch_size) + (1, 1)
    Fp, Hp, Wp = g[0], g[1], g[2]
    pt, ph, pw = p[0], p[1], p[2]
    if L != Fp * Hp * Wp or cols != pt * ph * pw * out_channels:
        raise _lib.MovaB200Error(f"unpatchify: x is {L}x{cols}, grid {g[:3]} patch {p[:3]} channels {out_channels}")
    out = torch.empty(out_channels, Fp * pt, Hp * ph, Wp * pw, dtype=torch.bfloat16, device=x.device)
    rc = _lib.load().mova_b200_unpatchify(x.data_ptr(), ldi, out.data_ptr(), out_channels, Fp, Hp, Wp, pt, ph, pw,
                                          _stream())
    _lib.check(rc, "mova_b200_unpatchify")
    return out.reshape(out_channels, Fp * pt) if one_d else out


def sinusoidal_embedding(dim: int, timestep: torch.Tensor) -> torch.Tensor:
    """sinusoidal_embedding_1d (wan_video_dit.py:99-103) for one timestep: fp32 ``[dim]``, fp64 math on the device;
    ``timestep`` is a one-element fp32 CUDA tensor and is not synchronised with the host."""
    _need(timestep, torch.float32, "timestep")
    if timestep.numel() != 1:
        raise _lib.MovaB200Error(f"sinusoidal_embedding: one timestep per call (MOVA runs B = 1), got {tuple(timestep.shape)}")
    out = torch.empty(dim, dtype=torch.float32, device=timestep.device)
    rc = _lib.load().mova_b200_sinusoidal(timestep.data_ptr(), out.data_ptr(), int(dim), _stream())
    _lib.check(rc, "mova_b200_sinusoidal")
    return out


def gemv_f32(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *, pre_silu: bool = False,
             post_silu: bool = False, want_bf16: bool = False):
    """``post(W pre(x) + b)`` for one fp32 activation vector and bf16 weights, fp32 accumulation and result: the
    M = 1 time_embedding / time_projection MLPs (wan_video_dit.py:374-380) as the reference evaluates them under
    autocast(float32) (pipeline_mova.py:544-549).  Returns fp32 ``[N]`` (and its bf16 rounding when ``want_bf16``)."""
    _need(x, torch.float32, "x")
    _need(weight, torch.bfloat16, "weight")
    N, K, ldw = _rows2d(weight, "weight")
    if x.numel() != K or not x.is_contiguous():
        raise _lib.MovaB200Error(f"gemv_f32: x must be a contiguous fp32 vector of K={K} elements, got {tuple(x.shape)}")
    bptr = None
    if bias is not None:
        _need(bias, torch.bfloat16, "bias")
        if bias.numel() != N or not bias.is_contiguous():
            raise _lib.MovaB200Error("gemv_f32: bias must be a contiguous bf16 vector of N elements")
        bptr = bias.data_ptr()
    y = torch.empty(N, dtype=torch.float32, device=x.device)
    yb = torch.empty(N, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    rc = _lib.load().mova_b200_gemv_f32(x.data_ptr(), weight.data_ptr(), ldw, bptr, y.data_ptr(),
                                        yb.data_ptr() if yb is not None else None, N, K, int(pre_silu), int(post_silu),
                                        _stream())
    _lib.check(rc, "mova_b200_gemv_f32")
    return (y, yb) if want_bf16 else y


def cfg_euler_step(posi: torch.Tensor, nega: Optional[torch.Tensor], sample: torch.Tensor, cfg_scale: float,
                   dsigma: float, *, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``sample + (nega + cfg_scale * (posi - nega)) * dsigma`` in one pass: the classifier-free-guidance combine of
    pipeline_mova.py:456-460 fused with ``FlowMatchPairScheduler.step_from_to`` (schedulers/flow_match_pair.py:213-227;
    ``dsigma = sigma_to - sigma_from`` from the reference's scheduler).  ``posi`` / ``nega``: bf16 model outputs
    (``nega=None``: ``cfg_scale == 1`` path, :439-441), ``sample``: fp32 latents; returns fp32 (``out`` may be ``sample``)."""
    _need(posi, torch.bfloat16, "posi")
    _need(sample, torch.float32, "sample")
    if posi.shape != sample.shape or not posi.is_contiguous() or not sample.is_contiguous():
        raise _lib.MovaB200Error(f"cfg_euler_step: posi {tuple(posi.shape)} and sample {tuple(sample.shape)} must be "
                                 "contiguous and of one shape")
    nptr = None
    if nega is not None:
        _need(nega, torch.bfloat16, "nega")
        if nega.shape != posi.shape or not nega.is_contiguous():
            raise _lib.MovaB200Error("cfg_euler_step: nega must match posi")
        nptr = nega.data_ptr()
    if out is None:
        out = torch.empty_like(sample)
    _need(out, torch.float32, "out")
    if out.shape != sample.shape or not out.is_contiguous():
        raise _lib.MovaB200Error("cfg_euler_step: bad `out`")
    rc = _lib.load().mova_b200_cfg_euler(posi.data_ptr(), nptr, sample.data_ptr(), out.data_ptr(), sample.numel(),
                                         float(cfg_scale), float(dsigma), _stream())
    _lib.check(rc, "mova_b200_cfg_euler")
    return out
