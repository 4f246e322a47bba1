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
