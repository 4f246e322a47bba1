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


def run(variant, *args):
    """(ok, TFLOP/s or None) of `selftest attn args` under MOVA_ATTN_VARIANT=variant; 60 s limit."""
    key = (variant,) + args
    if key in _cache:
        return _cache[key]
    if not os.path.exists(SELFTEST):
        pytest.skip("self-test binary not built (make -C dualforce_b200/csrc)")
    env = dict(os.environ, MOVA_ATTN_VARIANT=variant)
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
            f.write(json.dumps({"variant": variant, "args": args, "ok": ok, "tflops": res[1]}) + "\n")
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
