"""-m gpu: size-independent properties at the MOVA-720p geometry (BASELINE.json configs[3]: L_v = 49 x 45 x 80 =
176 400 video tokens) -- the largest sequence the path is specified for (green on B200 since round 1)."""
import pytest
import torch

from util import metrics

pytestmark = [pytest.mark.gpu]

S = 49 * 45 * 80


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


def test_attention_properties_at_720p_size():
    """One head of the 176 400-token self-attention (15.9 TFLOP): rows of the softmax sum to one (V = 1 gives 1),
    split-KV + LSE merge reproduces the unsplit result with a ragged split."""
    import dualforce_b200 as B

    ops = B.ops
    B._lib.require_device(0)
    q, k = rnd(1, S, 128, seed=1, scale=1.2).cuda(), rnd(1, S, 128, seed=2, scale=1.2).cuda()
    ones = torch.ones(1, S, 128, dtype=torch.bfloat16, device="cuda")
    o1 = ops.attention(q, k, ones, 1)
    assert (o1.float() - 1.0).abs().max() < 8e-3
    v = rnd(1, S, 128, seed=3).cuda()
    full = ops.attention(q, k, v, 1)
    cut = 100_003  # not a multiple of the 128-key block
    parts = [ops.attention(q, k[:, a:b], v[:, a:b], 1, return_lse=True) for a, b in ((0, cut), (cut, S))]
    merged = ops.lse_merge(torch.stack([o[0] for o, _ in parts]), torch.stack([l[0] for _, l in parts]), 1)
    m = metrics(merged, full[0])
    assert m["finite"] and m["rel_fro"] < 1e-2, m


def test_gemm_and_norms_at_720p_rows():
    """M = 176 400 rows: M x N exceeds 2^31 elements for the fused QKV projection, so every index on the way must be
    64-bit.  Checked on row samples against fp32 torch: QKV-shaped GEMM (N = 15360 would need 5.4 GB; N = 2560 with the
    same row count keeps the test light but still crosses 2^31 bytes), LayerNorm and RMSNorm + RoPE on the last rows."""
    import dualforce_b200 as B

    ops = B.ops
    d, N = 5120, 2560
    x = rnd(S, d, seed=4).cuda()
    w, b = rnd(N, d, seed=5, scale=d ** -0.5).cuda(), rnd(N, seed=6, scale=0.1).cuda()
    y = ops.linear(x, w, b)
    rows = torch.tensor([0, 1, 65535, 65536, 131071, 131072, S - 129, S - 1], device="cuda")
    ref = x[rows].float() @ w.float().t() + b.float()
    m = metrics(y[rows], ref)
    assert m["finite"] and m["ratio"] <= 6e-3 and m["rel_fro"] <= 4e-3, m
    ln = ops.layernorm(x, 1e-6)
    xr = x[rows].float()
    ref_ln = (xr - xr.mean(-1, keepdim=True)) * torch.rsqrt(xr.var(-1, unbiased=False, keepdim=True) + 1e-6)
    m = metrics(ln[rows], ref_ln)
    assert m["ratio"] <= 1e-2 and m["rel_fro"] <= 5e-3, m
    wn = torch.ones(d, dtype=torch.bfloat16, device="cuda")
    xn = x.clone()
    ops.rmsnorm_rope_(xn, wn, 1e-6)
    ref_rms = xr * torch.rsqrt((xr * xr).mean(-1, keepdim=True) + 1e-6)
    m = metrics(xn[rows], ref_rms)
    assert m["ratio"] <= 1e-2 and m["rel_fro"] <= 5e-3, m
