"""-m gpu: the experimental attention schedules (MOVA_ATTN_VARIANT=v7 / v8: Q.K^T of block j+1 issued in N-slices,
see csrc/attn.cu) against the shipped v3 schedule, through the device self-test binary (one process per variant: the
variant is latched at first use).  Opt-in kernels written after this round's GPU budget was spent -- the default path
never runs them -- so every test here is a non-strict xfail: XPASS/XFAIL is the first hardware verdict on each
variant (correct?  faster than v3?), nothing more.  The file sorts last so it cannot disturb the parity tests."""
import json
import os
import re
import subprocess

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="experimental schedules, first hardware run pending")]

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dualforce_b200", "csrc")
SELFTEST = os.path.join(CSRC, "selftest")
_cache = {}


def run(variant, *args, bounded=False):
    """(ok, TFLOP/s or None) of `selftest attn args` under MOVA_ATTN_VARIANT=variant (and MOVA_ATTN_BOUNDED=1);
    60 s limit."""
    key = (variant, bounded) + args
    if key in _cache:
        return _cache[key]
    if not os.path.exists(SELFTEST):
        pytest.skip("self-test binary not built (make -C dualforce_b200/csrc)")
    env = dict(os.environ, MOVA_ATTN_VARIANT=variant, MOVA_ATTN_BOUNDED="1" if bounded else "0")
    try:
        p = subprocess.run([SELFTEST, "attn", *map(str, args)], env=env, capture_output=True, text=True, timeout=60)
        out, ok = p.stdout, p.returncode == 0
    except subprocess.TimeoutExpired as exc:
        out, ok = (exc.stdout or b"").decode() if isinstance(exc.stdout, bytes) else (exc.stdout or ""), False
    m = re.search(r"([0-9.]+) TFLOP/s", out)
    res = (ok, float(m.group(1)) if m else None)
    _cache[key] = res
    try:  # best effort: leave the numbers where a gpurun call would collect them
        os.makedirs(os.path.join(os.path.dirname(CSRC), "..", "gpurun_out"), exist_ok=True)
        with open(os.path.join(os.path.dirname(CSRC), "..", "gpurun_out", "attn_variants.jsonl"), "a") as f:
            f.write(json.dumps({"variant": variant, "bounded": bounded, "args": args, "ok": ok, "tflops": res[1]}) + "\n")
    except OSError:
        pass
    return res


SHAPES = [(1, 128, 128, 1), (1, 256, 512, 2), (2, 300, 403, 3), (1, 403, 403, 12), (1, 403, 4400, 12), (1, 4400, 4400, 40)]


@pytest.mark.parametrize("variant", ["v7", "v8"])
def test_variant_matches_reference_kernel(variant):
    for shape in SHAPES:
        ok, _ = run(variant, *shape)
        assert ok, f"{variant} failed the self-test at B,Sq,Skv,H = {shape}"


@pytest.mark.parametrize("variant", ["v7", "v8"])
def test_variant_beats_v3_at_360p(variant):
    ok, tf = run(variant, 1, 43120, 43120, 40, 3)
    ok3, tf3 = run("v3", 1, 43120, 43120, 40, 3)
    assert ok and ok3 and tf is not None and tf3 is not None
    assert tf > 1.02 * tf3, f"{variant}: {tf} TFLOP/s vs v3 {tf3}"


def test_v8_beats_v7_at_360p():
    ok7, tf7 = run("v7", 1, 43120, 43120, 40, 3)
    ok8, tf8 = run("v8", 1, 43120, 43120, 40, 3)
    assert ok7 and ok8 and tf8 > tf7, f"v8 {tf8} vs v7 {tf7}"


# ---------------------------------------------------------------------------------------------- bounded softmax
@pytest.mark.parametrize("variant", ["v3", "v7", "v8"])
def test_bounded_softmax_matches_reference_kernel(variant):
    for shape in SHAPES:
        ok, _ = run(variant, *shape, bounded=True)
        assert ok, f"bounded {variant} failed the self-test at B,Sq,Skv,H = {shape}"


@pytest.mark.parametrize("variant", ["v3", "v8"])
def test_bounded_softmax_beats_v3_at_360p(variant):
    """Timing includes the two norm kernels (they run in every call of the self-test's bounded path)."""
    ok, tf = run(variant, 1, 43120, 43120, 40, 3, bounded=True)
    ok3, tf3 = run("v3", 1, 43120, 43120, 40, 3)
    assert ok and ok3 and tf is not None and tf3 is not None
    assert tf > 1.05 * tf3, f"bounded {variant}: {tf} TFLOP/s vs v3 {tf3}"


def test_bounded_attention_through_ops_both_paths():
    """ops.attention(bounded=True) vs the CPU oracle: unit-scale inputs (the bound holds: maxima are skipped) and
    inputs scaled until the Cauchy-Schwarz bound exceeds the slack (every block takes the exact path), plus the norms
    themselves."""
    import torch

    import dualforce_b200 as B
    import mova_oracle as O
    from util import assert_close

    B._lib.require_device(0)
    g = torch.Generator().manual_seed(0)
    H, Sq, Skv = 3, 300, 2500
    q = torch.randn(1, Sq, H * 128, generator=g).to(torch.bfloat16)
    k = torch.randn(1, Skv, H * 128, generator=g).to(torch.bfloat16)
    v = torch.randn(1, Skv, H * 128, generator=g).to(torch.bfloat16)
    rn, _ = B.ops.head_norms(q.cuda(), H, rows=True)
    _, bm = B.ops.head_norms(k.cuda(), H, blocks=True)
    ref_rn = q.float().reshape(1, Sq, H, 128).norm(dim=-1)
    assert (rn.cpu() - ref_rn).abs().max() <= 1e-4 * ref_rn.max()
    kn = k.float().reshape(1, Skv, H, 128).norm(dim=-1)
    ref_bm = torch.stack([kn[:, s:s + 128].amax(dim=1) for s in range(0, Skv, 128)], dim=-1)  # [1, H, nblk]
    assert bm.shape == ref_bm.shape and (bm.cpu() - ref_bm).abs().max() <= 1e-4 * ref_bm.max()
    for qscale in (1.0, 6.0):  # 6x: |q||k| * scale * log2(e) ~ 100 > slack 64 -> exact path
        qs = (q.float() * qscale).to(torch.bfloat16)
        got, lse = B.ops.attention(qs.cuda(), k.cuda(), v.cuda(), H, return_lse=True, bounded=True)
        ref, ref_lse = O.attention(qs.float(), k.float(), v.float(), H, return_lse=True)
        assert_close(got, ref, f"bounded attention, q x{qscale}", ratio=1.5e-2, fro=1e-2)
        assert (lse.cpu() - ref_lse).abs().max() <= 2e-3 * max(1.0, ref_lse.abs().max().item())
        plain = B.ops.attention(qs.cuda(), k.cuda(), v.cuda(), H, bounded=False)
        assert_close(got, plain.float().cpu(), f"bounded vs plain kernel, q x{qscale}", ratio=1.5e-2, fro=6e-3)


def test_split_kv_bridge_attention_on_device(monkeypatch):
    """MOVA_V2A_SPLITS=5 at the real v2a shape (403 audio queries, 43 120 video keys, 12 heads): same result as the
    single-launch path, and faster (24 -> 120 CTAs)."""
    import torch

    import dualforce_b200 as B
    from util import assert_close

    B._lib.require_device(0)
    g = torch.Generator().manual_seed(2)
    cca = B.ConditionalCrossAttention(1536, 5120, 12)
    for prm in cca.parameters():
        prm.data = torch.randn(prm.shape, generator=g) * (0.02 if prm.dim() > 1 else 0.1)
    cca.to("cuda", torch.bfloat16)
    x = torch.randn(1, 403, 1536, generator=g).to(torch.bfloat16).cuda()
    y = torch.randn(1, 43120, 5120, generator=g).to(torch.bfloat16).cuda()

    def timed(n):
        monkeypatch.setenv("MOVA_V2A_SPLITS", str(n))
        out = cca.attend(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            cca.attend(x, y)
        e1.record()
        e1.synchronize()
        return out, e0.elapsed_time(e1) / 5

    plain, t1 = timed(1)
    split, t5 = timed(5)
    assert_close(split, plain.float().cpu(), "split-KV vs unsplit", ratio=1e-2, fro=6e-3)
    assert t5 < t1, f"split {t5:.3f} ms vs unsplit {t1:.3f} ms (projections included in both)"
