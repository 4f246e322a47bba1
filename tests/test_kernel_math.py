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


def test_bounded_softmax_overflow_budget():
    """P <= 2^slack per element; the row sum over the longest specified sequence and the O accumulator (|v| up to
    2^8) must stay far inside fp32 (2^127), and the skip test must be conservative: bound >= any score."""
    src = open(os.path.join(CSRC, "attn.cu")).read()
    slack = float(re.search(r"AT_BOUND_SLACK = ([0-9.]+)f", src).group(1))
    longest = 49 * 45 * 80  # 720p
    assert slack + math.log2(longest) + 8 < 120
    rng = np.random.default_rng(0)
    q = rng.standard_normal((64, 128)).astype(np.float32)
    k = rng.standard_normal((128, 128)).astype(np.float32) * 3
    bound = np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(k, axis=1).max()
    assert (q @ k.T <= bound * 1.001).all()


def _bf16(x):
    """Round-to-nearest-even to bfloat16, kept in float32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32)
    r = ((u.astype(np.uint64) + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def _online_softmax(q, k, v, scale, bounded, threshold=8.0, slack=64.0):
    """Tile-level model of csrc/attn.cu's softmax in float32: exp2 domain, lazy rescale, P rounded to bf16 before P.V,
    optional bounded skip of the row maximum (reference point left stale while the Cauchy-Schwarz bound allows)."""
    c = np.float32(scale * 1.4426950408889634)
    Sq, D = q.shape
    o = np.zeros((Sq, D), np.float32)
    l = np.zeros(Sq, np.float32)
    m_used = np.full(Sq, -np.inf, np.float32)
    qn_c = np.linalg.norm(q, axis=1).astype(np.float32) * c * np.float32(1.001)
    skipped = 0
    for j0 in range(0, k.shape[0], 128):
        kb, vb = k[j0:j0 + 128], v[j0:j0 + 128]
        s = (q @ kb.T).astype(np.float32)
        kmax = np.float32(np.linalg.norm(kb, axis=1).max())
        skip = bounded and j0 > 0 and bool(np.all(qn_c * kmax - m_used * c <= slack))
        if skip:
            skipped += 1
        else:
            m_new = np.maximum(m_used, s.max(axis=1))
            if j0 == 0:
                m_used = m_new
            elif np.any((m_new - m_used) * c > threshold):
                f = np.exp2((m_used - m_new) * c).astype(np.float32)
                m_used, l, o = m_new, l * f, o * f[:, None]
        p = np.exp2(s * c - (m_used * c)[:, None]).astype(np.float32)
        l = l + p.sum(axis=1, dtype=np.float32)
        o = o + _bf16(p) @ vb
    return o / l[:, None], m_used * np.float32(scale) + np.log(l), skipped


def test_bounded_softmax_is_numerically_equivalent():
    """The stale-reference-point softmax (bounded skip) against exact attention in float64 and against the same model
    with the maximum taken in every block: same accuracy, although P grows to 2^20..2^40 on the way."""
    rng = np.random.default_rng(3)
    Sq, Skv, D = 64, 4096, 128
    q = _bf16(rng.standard_normal((Sq, D)))
    k = _bf16(rng.standard_normal((Skv, D)))
    k[2000:2010] *= 2.5  # late, much larger scores: the reference point of block 0 goes stale by ~2^30
    v = _bf16(rng.standard_normal((Skv, D)))
    scale = 3.0 / math.sqrt(D)  # the self-test's sharpened scale
    s = (q.astype(np.float64) @ k.astype(np.float64).T) * scale
    pexact = np.exp(s - s.max(axis=1, keepdims=True))
    exact = (pexact / pexact.sum(axis=1, keepdims=True)) @ v.astype(np.float64)
    lse_exact = s.max(axis=1) + np.log(pexact.sum(axis=1))
    o_plain, lse_plain, skipped_plain = _online_softmax(q, k, v, scale, bounded=False)
    o_bound, lse_bound, skipped = _online_softmax(q, k, v, scale, bounded=True)
    assert skipped_plain == 0 and skipped >= 28  # nearly every block after the first skips its maximum
    err_plain = np.abs(o_plain - exact).max()
    err_bound = np.abs(o_bound - exact).max()
    assert err_bound <= 1.5 * err_plain + 1e-6 and err_bound < 1e-2, (err_plain, err_bound)  # bf16 rounding of P
    assert np.abs(lse_bound - lse_exact).max() < 1e-3 and np.abs(lse_plain - lse_exact).max() < 1e-3
    # inputs for which the bound is useless (huge norms): every block takes the exact path, same answer as plain
    o_big, _, skipped_big = _online_softmax(q * 8, k, v, scale, bounded=True)
    o_big_plain, _, _ = _online_softmax(q * 8, k, v, scale, bounded=False)
    assert skipped_big == 0 and np.array_equal(o_big, o_big_plain)
