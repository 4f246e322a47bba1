"""TEST INFRASTRUCTURE ONLY -- a torch-CPU stand-in for ``dualforce_b200.ops`` (i.e. for ``libmova_b200.so``).

The product has no CPU path; without a GPU none of its Python host logic (module twins, weight packing, view /
stride plumbing, the dual-tower loop, the step wrapper and its memo caches) could be exercised by the ``-m "not gpu"``
suite.  ``install(monkeypatch)`` swaps every entry point of ``dualforce_b200.ops`` for a function with the same
signature and the same storage semantics (bf16 in / bf16 out, fp32 math, in-place where the kernel is in place,
strided views, segmented context-parallel layouts), so the host code runs unchanged on CPU tensors and can be
checked against the oracle.  Nothing under ``dualforce_b200/`` imports this file (tests/test_abi.py checks), and the
index arithmetic of patchify / unpatchify below is a literal transcription of the CUDA kernels' (csrc/step.cu), so
the CPU suite also checks that arithmetic against the oracle's einops-style restatement.
"""
from __future__ import annotations

import math

import torch

BF16, F32 = torch.bfloat16, torch.float32
EPI_BIAS, EPI_GELU_TANH, EPI_RESIDUAL = 0, 1, 2
ROPE_NONE, ROPE_INTERLEAVED, ROPE_HALF = 0, 1, 2

CALLS = {}  # entry point -> number of calls (the tests assert on launch counts / cache hits)


def _count(name):
    CALLS[name] = CALLS.get(name, 0) + 1


def _need(t, dtype, name):
    assert isinstance(t, torch.Tensor) and t.dtype == dtype, f"{name}: expected {dtype}, got {getattr(t, 'dtype', type(t))}"


def _gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def linear(x, weight, bias=None, *, epilogue=EPI_BIAS, residual=None, gate=None, scale=1.0, out=None, cta_group=0,
           segments=1, out_segments=1):
    _count("linear")
    _need(x, BF16, "x"); _need(weight, BF16, "weight")
    if segments > 1:
        assert x.dim() == 3 and x.shape[0] == segments and x.is_contiguous()
        M = x.shape[1]
        a = x.permute(1, 0, 2).reshape(M, -1).float()
        out_shape = (M,)
    else:
        assert x.stride(-1) == 1
        a = x.reshape(-1, x.shape[-1]).float()
        M = a.shape[0]
        out_shape = tuple(x.shape[:-1])
    N, K = weight.shape
    assert a.shape[1] == K and N % 8 == 0 and K % 8 == 0
    c = a @ weight.float().t()
    if bias is not None:
        _need(bias, BF16, "bias")
        c = c + bias.float()
    if epilogue == EPI_GELU_TANH:
        c = _gelu_tanh(c)
    elif epilogue == EPI_RESIDUAL:
        assert residual is not None and out_segments == 1
        _need(residual, BF16, "residual")
        g = torch.ones(N) if gate is None else gate
        if gate is not None:
            _need(gate, F32, "gate")
            assert gate.numel() == N and gate.is_contiguous()
        c = residual.reshape(-1, N).float() + g * float(scale) * c
    c = c.to(BF16)
    if out_segments > 1:
        seg_n = N // out_segments
        if out is None:
            out = torch.empty(out_segments, M, seg_n, dtype=BF16)
        assert tuple(out.shape) == (out_segments, M, seg_n) and out.is_contiguous()
        out.copy_(c.reshape(M, out_segments, seg_n).permute(1, 0, 2))
        return out
    if out is None:
        return c.reshape(*out_shape, N)
    _need(out, BF16, "out")
    assert out.numel() == M * N and out.stride(-1) == 1
    out.copy_(c.reshape(out.shape))
    return out


def attention(q, k, v, num_heads, *, return_lse=False, softmax_scale=None, out=None, variant=None, emu=4):
    _count("attention")
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _need(t, BF16, n)
        assert t.dim() == 3 and t.stride(2) == 1
    B, Sq, HD = q.shape
    D = HD // num_heads
    assert D == 128, "the sm_100a kernel is head_dim 128 only"
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(D)
    qh = q.float().reshape(B, Sq, num_heads, D).permute(0, 2, 1, 3)
    kh = k.float().reshape(B, -1, num_heads, D).permute(0, 2, 1, 3)
    vh = v.float().reshape(B, -1, num_heads, D).permute(0, 2, 1, 3)
    s = (qh @ kh.transpose(-1, -2)) * scale
    lse = torch.logsumexp(s, dim=-1)
    o = (torch.exp(s - lse.unsqueeze(-1)) @ vh).permute(0, 2, 1, 3).reshape(B, Sq, HD).to(BF16)
    if out is not None:
        out.copy_(o)
        o = out
    return (o, lse.contiguous()) if return_lse else o


def layernorm(x, eps, *, weight=None, bias=None, shift=None, scale=None, out=None):
    _count("layernorm")
    _need(x, BF16, "x")
    d = x.shape[-1]
    assert d % 8 == 0 and (weight is None) == (bias is None) and (shift is None) == (scale is None)
    xf = x.float()
    mu = xf.mean(-1, keepdim=True)
    var = ((xf - mu) ** 2).mean(-1, keepdim=True)
    y = (xf - mu) * torch.rsqrt(var + eps)
    if weight is not None:
        _need(weight, BF16, "weight"); _need(bias, BF16, "bias")
        y = y * weight.float() + bias.float()
    if shift is not None:
        _need(shift, F32, "shift"); _need(scale, F32, "scale")
        assert shift.numel() == d and shift.is_contiguous() and scale.numel() == d and scale.is_contiguous()
        y = y * (1.0 + scale) + shift
    y = y.to(BF16)
    if out is None:
        return y
    _need(out, BF16, "out")
    out.copy_(y.reshape(out.shape))
    return out


def rmsnorm_rope_(x, weight, eps, *, head_dim=128, cos=None, sin=None, rope_mode=ROPE_NONE, segments=1, seg_stride=0):
    if x.dim() == 3 and x.shape[0] > 1:  # a batch sharing one position table (CFG pair): sample by sample, like ops.py
        for b in range(x.shape[0]):
            rmsnorm_rope_(x[b], weight, eps, head_dim=head_dim, cos=cos, sin=sin, rope_mode=rope_mode, segments=segments,
                          seg_stride=seg_stride)
        return x
    _count("rmsnorm_rope_")
    _need(x, BF16, "x"); _need(weight, BF16, "weight")
    assert x.stride(-1) == 1
    seg_len = x.shape[-1]
    L = x.numel() // seg_len
    row_stride = x.stride(-2)
    for dim in range(x.dim() - 2):
        assert x.shape[dim] == 1, "emulation handles [1, L, d] / [L, d] views"
    d = seg_len * segments
    assert d % 128 == 0 and weight.numel() == d
    view = torch.as_strided(x, (segments, L, seg_len), (int(seg_stride), row_stride, 1), x.storage_offset())
    full = view.permute(1, 0, 2).reshape(L, d).float()
    y = full * torch.rsqrt((full * full).mean(-1, keepdim=True) + eps) * weight.float()
    if rope_mode != ROPE_NONE:
        assert head_dim == 128
        _need(cos, F32, "cos"); _need(sin, F32, "sin")
        H = d // head_dim
        yh = y.reshape(L, H, head_dim)
        if rope_mode == ROPE_INTERLEAVED:
            assert cos.shape == (L, 64) and sin.shape == (L, 64) and cos.is_contiguous() and sin.is_contiguous()
            re, im = yh[..., 0::2], yh[..., 1::2]
            c, s = cos[:, None, :], sin[:, None, :]
            yh = torch.stack([re * c - im * s, re * s + im * c], dim=-1).reshape(L, H, head_dim)
        else:
            assert cos.shape == (L, 128) and sin.shape == (L, 128) and cos.is_contiguous() and sin.is_contiguous()
            x1, x2 = yh[..., :64], yh[..., 64:]
            rot = torch.cat((-x2, x1), dim=-1)
            yh = yh * cos[:, None, :] + rot * sin[:, None, :]
        y = yh.reshape(L, d)
    view.copy_(y.to(BF16).reshape(L, segments, seg_len).permute(1, 0, 2))
    return x


def lse_merge(o_parts, lse_parts, num_heads, *, return_lse=False):
    _count("lse_merge")
    _need(o_parts, BF16, "o_parts"); _need(lse_parts, F32, "lse_parts")
    P, rows, HD = o_parts.shape
    D = HD // num_heads
    assert lse_parts.shape == (P, num_heads, rows)
    lse = torch.logsumexp(lse_parts, dim=0)  # [H, rows]
    w = torch.exp(lse_parts - lse)  # [P, H, rows]
    o = (o_parts.float().reshape(P, rows, num_heads, D) * w.permute(0, 2, 1).unsqueeze(-1)).sum(0).reshape(rows, HD)
    o = o.to(BF16)
    return (o, lse) if return_lse else o


def add_to_f32(a, b=None):
    _count("add_to_f32")
    _need(a, BF16, "a")
    out = a.float()
    if b is not None:
        _need(b, BF16, "b")
        out = out + b.expand_as(a).float()
    return out.contiguous()


def patchify(x, patch_size, *, out=None):
    _count("patchify")
    assert x.dtype in (F32, BF16)
    if x.dim() == 2:
        x = x[:, :, None, None]
    assert x.dim() == 4 and x.is_contiguous()
    p = tuple(int(v) for v in patch_size) + (1, 1)
    pt, ph, pw = p[0], p[1], p[2]
    C, F, H, W = x.shape
    assert F % pt == 0 and H % ph == 0 and W % pw == 0
    Hp, Wp = H // ph, W // pw
    K = C * pt * ph * pw
    L = (F // pt) * Hp * Wp
    # literal transcription of patchify_kernel (csrc/step.cu): one "thread" per output element
    idx = torch.arange(L * K, dtype=torch.int64)
    l = idx // K
    k = idx - l * K
    w = l % Wp
    h = (l // Wp) % Hp
    f = l // (Wp * Hp)
    dw = k % pw; k = k // pw
    dh = k % ph; k = k // ph
    dt = k % pt
    c = k // pt
    src = ((c * F + (f * pt + dt)) * H + (h * ph + dh)) * W + (w * pw + dw)
    cols = x.reshape(-1)[src].float().to(BF16).reshape(L, K)
    if out is not None:
        out.copy_(cols)
        return out
    return cols


def unpatchify(x, grid_size, patch_size, out_channels):
    _count("unpatchify")
    _need(x, BF16, "x")
    g = tuple(int(v) for v in grid_size)
    one_d = len(g) == 1
    g = g + (1, 1)
    p = tuple(int(v) for v in patch_size) + (1, 1)
    Fp, Hp, Wp = g[0], g[1], g[2]
    pt, ph, pw = p[0], p[1], p[2]
    L, cols = x.shape
    assert L == Fp * Hp * Wp and cols == pt * ph * pw * out_channels and x.stride(1) == 1
    ldi = x.stride(0)
    Wo, Ho, Fo = Wp * pw, Hp * ph, Fp * pt
    # literal transcription of unpatchify_kernel (csrc/step.cu)
    idx = torch.arange(out_channels * Fo * Ho * Wo, dtype=torch.int64)
    wo = idx % Wo
    ho = (idx // Wo) % Ho
    fo = (idx // (Wo * Ho)) % Fo
    c = idx // (Wo * Ho * Fo)
    w, z = wo // pw, wo % pw
    h, y = ho // ph, ho % ph
    f, xx = fo // pt, fo % pt
    l = (f * Hp + h) * Wp + w
    col = ((xx * ph + y) * pw + z) * out_channels + c
    span = (L - 1) * ldi + cols  # elements from the first to one past the last addressed one
    flat = torch.as_strided(x, (span,), (1,), x.storage_offset())
    out = flat[l * ldi + col].reshape(out_channels, Fo, Ho, Wo)
    return out.reshape(out_channels, Fo) if one_d else out


def sinusoidal_embedding(dim, timestep):
    _count("sinusoidal_embedding")
    _need(timestep, F32, "timestep")
    assert timestep.numel() == 1 and dim % 2 == 0
    half = dim // 2
    pos = timestep.reshape(()).to(torch.float64)
    ang = pos * torch.pow(torch.tensor(10000.0, dtype=torch.float64),
                          -torch.arange(half, dtype=torch.float64) / half)
    return torch.cat([torch.cos(ang), torch.sin(ang)]).to(F32)


def _silu(v):
    return v / (1.0 + torch.exp(-v))


def gemv_f32(x, weight, bias=None, *, pre_silu=False, post_silu=False, want_bf16=False):
    _count("gemv_f32")
    _need(x, F32, "x"); _need(weight, BF16, "weight")
    N, K = weight.shape
    assert x.numel() == K and x.is_contiguous() and K % 8 == 0
    xv = _silu(x.reshape(-1)) if pre_silu else x.reshape(-1)
    y = weight.float() @ xv
    if bias is not None:
        _need(bias, BF16, "bias")
        y = y + bias.float()
    if post_silu:
        y = _silu(y)
    return (y, y.to(BF16)) if want_bf16 else y


def cfg_euler_step(posi, nega, sample, cfg_scale, dsigma, *, out=None):
    _count("cfg_euler_step")
    _need(posi, BF16, "posi"); _need(sample, F32, "sample")
    g = posi.float()
    if nega is not None:
        _need(nega, BF16, "nega")
        g = nega.float() + float(cfg_scale) * (posi.float() - nega.float())
    res = sample + g * float(dsigma)
    if out is not None:
        out.copy_(res)
        return out
    return res


ENTRY_POINTS = dict(linear=linear, attention=attention, layernorm=layernorm, rmsnorm_rope_=rmsnorm_rope_,
                    lse_merge=lse_merge, add_to_f32=add_to_f32, patchify=patchify, unpatchify=unpatchify,
                    sinusoidal_embedding=sinusoidal_embedding, gemv_f32=gemv_f32,
                    cfg_euler_step=cfg_euler_step)


def install(monkeypatch):
    """Swap the emulation in for the duration of one test (pytest's monkeypatch restores the real bindings)."""
    import dualforce_b200.ops as real

    missing = [n for n in real.__all__ if not n.startswith(("EPI_", "ROPE_")) and n not in ENTRY_POINTS]
    assert not missing, f"emulated_ops is missing entry points: {missing}"
    for name, fn in ENTRY_POINTS.items():
        monkeypatch.setattr(real, name, fn)
    CALLS.clear()
    return CALLS
