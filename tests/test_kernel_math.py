"""Closed-form pieces of the device code, re-evaluated on the CPU with the same constants (parsed out of the .cu / .cuh
sources, so the test follows the kernels): the degree-3 exp2 polynomial of the attention softmax, the sigmoid form of
GELU-tanh in the GEMM epilogue, and the bounded-softmax overflow budget."""
import math
import os
import re

import numpy as np

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dualforce_b200", "csrc")


def test_exp2_polynomial_error_bound():
    src = open(os.path.join(CSRC, "attn_common.cuh")).read()
    body = src[src.index("exp2_poly2(float2 x)"):]
    coeffs = [float(c) for c in re.findall(r"make_float2\(([0-9.]+)f, \1f\)", body)[:5]]
    # magic, then c3, c2, c1, c0 in Horner order
    assert coeffs[0] == 12582912.0 and len(coeffs) == 5
    c3, c2, c1, c0 = coeffs[1:]
    x = np.linspace(-30.0, 30.0, 2_000_001, dtype=np.float64)
    n = np.rint(x)
    r = (x - n).astype(np.float32)
    p = ((np.float32(c3) * r + np.float32(c2)) * r + np.float32(c1)) * r + np.float32(c0)
    approx = p.astype(np.float64) * np.exp2(n)
    rel = np.abs(approx - np.exp2(x)) / np.exp2(x)
    assert rel.max() < 1.0e-4, rel.max()  # header claims 7.5e-5; bf16 rounding of P is 3.9e-3


def test_gelu_tanh_sigmoid_form():
    src = open(os.path.join(CSRC, "gemm.cu")).read()
    assert "0.7978845608028654f" in src and "0.044715f" in src
    x = np.linspace(-12, 12, 200001)
    u = 0.7978845608028654 * x * (1.0 + 0.044715 * x * x)
    sig = x / (1.0 + np.exp2(-2.0 * 1.4426950408889634 * u))
    ref = 0.5 * x * (1.0 + np.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))
    assert np.abs(sig - ref).max() < 1e-12


def _rescale_thresholds():
    out = {}
    for fn, name in (("attn.cu", "AT_RESCALE_THRESHOLD"), ("attn_pair.cu", "AP_RESCALE_THRESHOLD")):
        src = open(os.path.join(CSRC, fn)).read()
        out[fn] = float(re.search(name + r" = ([0-9.]+)f", src).group(1))
    return out


def test_lazy_rescale_overflow_budget():
    """Both attention kernels only advance the reference maximum when a block maximum exceeds it by more than
    2^threshold (log2 units): P <= 2^threshold per element, so the row sum over the longest specified sequence and the
    O accumulator (|v| up to 2^8) must stay far inside fp32 (2^127) and bf16 P must not overflow (2^127 as well)."""
    th = _rescale_thresholds()
    assert th["attn.cu"] == th["attn_pair.cu"], th
    longest = 49 * 45 * 80  # 720p
    for t in th.values():
        assert t + math.log2(longest) + 8 < 120


def _bf16(x):
    """Round-to-nearest-even to bfloat16, kept in float32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32)
    r = ((u.astype(np.uint64) + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def _online_softmax(q, k, v, scale, threshold, block=128):
    """Tile-level model of the kernels' softmax in float32: exp2 domain, reference maximum advanced only past
    ``threshold`` (0 = every block, the textbook form), P rounded to bf16 before P.V, row sum from the unrounded P."""
    c = np.float32(scale * 1.4426950408889634)
    Sq, D = q.shape
    o = np.zeros((Sq, D), np.float32)
    l = np.zeros(Sq, np.float32)
    m_used = np.full(Sq, -np.inf, np.float32)
    rescales = 0
    for j0 in range(0, k.shape[0], block):
        kb, vb = k[j0:j0 + block], v[j0:j0 + block]
        s = (q @ kb.T).astype(np.float32)
        m_new = np.maximum(m_used, s.max(axis=1))
        if j0 == 0:
            m_used = m_new
        elif np.any((m_new - m_used) * c > threshold):  # warp-uniform vote in the kernels
            f = np.exp2((m_used - m_new) * c).astype(np.float32)
            m_used, l, o = m_new, l * f, o * f[:, None]
            rescales += 1
        p = np.exp2(s * c - (m_used * c)[:, None]).astype(np.float32)
        l = l + p.sum(axis=1, dtype=np.float32)
        o = o + _bf16(p) @ vb
    return o / l[:, None], m_used * np.float32(scale) + np.log(l), rescales


def test_lazy_rescale_softmax_is_numerically_equivalent():
    """The lazily advanced reference point against exact attention in float64 and against the same model with the
    maximum advanced in every block: same accuracy with far fewer accumulator rescales, including when late keys carry
    much larger scores than the first block."""
    th = _rescale_thresholds()["attn_pair.cu"]
    rng = np.random.default_rng(3)
    Sq, Skv, D = 64, 4096, 128
    q = _bf16(rng.standard_normal((Sq, D)))
    k = _bf16(rng.standard_normal((Skv, D)))
    k[2000:2010] *= 2.5  # late, much larger scores
    v = _bf16(rng.standard_normal((Skv, D)))
    scale = 3.0 / math.sqrt(D)  # the self-test's sharpened scale
    s = (q.astype(np.float64) @ k.astype(np.float64).T) * scale
    pexact = np.exp(s - s.max(axis=1, keepdims=True))
    exact = (pexact / pexact.sum(axis=1, keepdims=True)) @ v.astype(np.float64)
    lse_exact = s.max(axis=1) + np.log(pexact.sum(axis=1))
    o_every, lse_every, n_every = _online_softmax(q, k, v, scale, threshold=0.0)
    o_lazy, lse_lazy, n_lazy = _online_softmax(q, k, v, scale, threshold=th)
    assert n_lazy < n_every and n_lazy >= 1  # the late keys force at least one real rescale
    err_every = np.abs(o_every - exact).max()
    err_lazy = np.abs(o_lazy - exact).max()
    # both are bf16-rounding-of-P noise (relative 2^-9 whatever the reference point): same order of magnitude
    assert err_lazy <= 4 * err_every + 1e-6 and err_lazy < 1e-2, (err_every, err_lazy)
    assert np.abs(lse_lazy - lse_exact).max() < 1e-3 and np.abs(lse_every - lse_exact).max() < 1e-3
