#!/usr/bin/env python
"""BASELINE.json configs[4]: bidirectional video<->audio cross-attention microbench -- video token count x audio
token count x heads, this repo's attention kernel against the reference's attention (`flash_attention()`,
mova/diffusion/models/wan_video_dit.py:58-91: flash_attn when importable, else torch SDPA) on the same tensors.

    python benchmarks/bridge_microbench.py [--quick] [--iters 20]      # one JSON line per case, on one B200

a2v: queries = video tokens (H = 40 heads), keys = audio tokens;  v2a: queries = audio tokens (H = 12), keys = video
tokens.  Per-rank head counts of the context-parallel path (20 / 10 / 5) are included for a2v.  Timing: CUDA events,
3 warm-up + `iters` launches, L2 flushed between launches by a 256 MB memset; parity of the two implementations is
checked on every case (bf16 tolerance)."""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def reference_attention(q, k, v, num_heads):
    """The reference's flash_attention() dispatch, restated: FA2 if importable, else SDPA (wan_video_dit.py:70-90)."""
    B, Sq, HD = q.shape
    D = HD // num_heads
    try:
        import flash_attn

        qh, kh, vh = (t.reshape(B, -1, num_heads, D) for t in (q, k, v))
        return flash_attn.flash_attn_func(qh, kh, vh).reshape(B, Sq, HD), "flash_attn"
    except Exception:
        qh, kh, vh = (t.reshape(B, -1, num_heads, D).transpose(1, 2) for t in (q, k, v))
        o = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)
        return o.transpose(1, 2).reshape(B, Sq, HD), "sdpa"


def sdpa_cudnn(q, k, v, num_heads):
    """torch SDPA with the cuDNN fused-attention back end forced (the fastest library attention on this box,
    profiles/r02_kernels_vs_libs*.jsonl) -- the bar beside the reference's own choice."""
    from torch.nn.attention import SDPBackend, sdpa_kernel

    B, Sq, HD = q.shape
    qh, kh, vh = (t.view(B, -1, num_heads, HD // num_heads).transpose(1, 2) for t in (q, k, v))
    with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
        return torch.nn.functional.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, Sq, HD)


def time_fn(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import dualforce_b200 as B

    B._lib.require_device(0)
    lvs = [4400, 43120] if args.quick else [4400, 10780, 21560, 43120, 176400]
    las = [403] if args.quick else [101, 202, 403, 806, 1612]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    for lv in lvs:
        for la in las:
            for name, sq, skv, heads in (("a2v", lv, la, 40), ("a2v/cp2", lv, la, 20), ("a2v/cp8", lv, la, 5),
                                         ("v2a", la, lv, 12)):
                q = torch.randn(1, sq, heads * 128, device="cuda", generator=g).to(torch.bfloat16)
                k = torch.randn(1, skv, heads * 128, device="cuda", generator=g).to(torch.bfloat16)
                v = torch.randn(1, skv, heads * 128, device="cuda", generator=g).to(torch.bfloat16)
                attn = B.AttentionModule(heads)  # the product's attention processor (splits the keys for v2a)
                ours = attn(q, k, v)
                ref, backend = reference_attention(q, k, v, heads)
                err = (ours.float() - ref.float()).abs().max().item() / max(ref.float().abs().max().item(), 1e-30)
                t_ours = time_fn(lambda: attn(q, k, v), args.iters, flush)
                t_ref = time_fn(lambda: reference_attention(q, k, v, heads), args.iters, flush)
                flops = 4.0 * heads * sq * skv * 128
                try:
                    t_cudnn = time_fn(lambda: sdpa_cudnn(q, k, v, heads), args.iters, flush)
                except Exception:  # noqa: BLE001 -- back end unavailable for this shape
                    t_cudnn = None
                print(json.dumps({"case": name, "L_v": lv, "L_a": la, "heads": heads, "ms": t_ours, "ref_ms": t_ref,
                                  "ref_backend": backend, "speedup": t_ref / t_ours, "sdpa_cudnn_ms": t_cudnn,
                                  "speedup_vs_sdpa_cudnn": (t_cudnn / t_ours) if t_cudnn else None,
                                  "tflops": flops / t_ours * 1e-9, "ref_tflops": flops / t_ref * 1e-9,
                                  "max_err_ratio": err, "parity_ok": bool(err < 2e-2 and math.isfinite(err))}),
                      flush=True)
                del q, k, v, ours, ref


if __name__ == "__main__":
    main()
