#!/usr/bin/env python
"""Same-box bar: this repo's tcgen05 attention and GEMM kernels timed next to the libraries the
reference would run on a B200 -- flash_attn (the reference's first choice, wan_video_dit.py:70-78), torch SDPA with
the cuDNN and the flash back ends forced separately (:85-90) and cuBLASLt through F.linear -- in one process, on the
MOVA-360p shapes.

    python benchmarks/kernels_vs_libs.py [--iters 5] [--quick] > profiles/r02_kernels_vs_libs.jsonl

One JSON line per (op, shape, implementation).  Timing: CUDA events around each launch, 3 warm-up launches, a 256 MB
memset between timed launches (L2 flush); the median is reported.  Clocks are sampled by the caller
(benchmarks/gpu_call.sh)."""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def time_fn(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def attention_cases(args, B, flush):
    from torch.nn.attention import SDPBackend, sdpa_kernel

    try:
        import flash_attn
    except Exception as e:  # noqa: BLE001
        flash_attn = None
        emit(op="attention", impl="flash_attn", error=f"import failed: {e!r}")
    cases = [(43120, 43120, 40), (43120, 43120, 5), (43120, 512, 40), (43120, 403, 40), (403, 43120, 12)]
    if not args.quick:
        cases[1:1] = [(43120, 43120, 20), (43120, 43120, 10)]
    g = torch.Generator(device="cuda").manual_seed(0)
    for sq, skv, h in cases:
        q = torch.randn(1, sq, h * 128, device="cuda", generator=g).to(torch.bfloat16)
        k = torch.randn(1, skv, h * 128, device="cuda", generator=g).to(torch.bfloat16)
        v = torch.randn(1, skv, h * 128, device="cuda", generator=g).to(torch.bfloat16)
        flops = 4.0 * h * sq * skv * 128
        qh, kh, vh = (t.view(1, -1, h, 128) for t in (q, k, v))
        qt, kt, vt = (t.transpose(1, 2) for t in (qh, kh, vh))
        impls = {"dualforce_b200": lambda: B.ops.attention(q, k, v, h)}
        if flash_attn is not None:
            impls["flash_attn"] = lambda: flash_attn.flash_attn_func(qh, kh, vh).reshape(1, sq, h * 128)

        def sdpa(backend):
            def run():
                with sdpa_kernel(backend):
                    return F.scaled_dot_product_attention(qt, kt, vt).transpose(1, 2).reshape(1, sq, h * 128)
            return run

        impls["sdpa_cudnn"] = sdpa(SDPBackend.CUDNN_ATTENTION)
        impls["sdpa_flash"] = sdpa(SDPBackend.FLASH_ATTENTION)
        ref = None
        for name, fn in impls.items():
            try:
                out = fn()
                torch.cuda.synchronize()
                if ref is None:
                    ref = out.float()
                    err = 0.0
                else:
                    err = (out.float() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
                med, best = time_fn(fn, args.iters, flush)
                emit(op="attention", Sq=sq, Skv=skv, H=h, impl=name, ms=med, ms_best=best,
                     tflops=flops / med * 1e-9, tflops_best=flops / best * 1e-9, max_err_vs_ours=err)
            except Exception as e:  # noqa: BLE001
                emit(op="attention", Sq=sq, Skv=skv, H=h, impl=name, error=repr(e)[:300])
        del q, k, v, ref


def gemm_cases(args, B, flush):
    # (name, M, N, K): the video-tower linears of one DiTBlock / bridge layer at 360p, cp = 1 and cp = 8
    shapes = [("qkv", 43120, 15360, 5120), ("o_proj", 43120, 5120, 5120), ("ffn1", 43120, 13824, 5120),
              ("ffn2", 43120, 5120, 13824), ("v2a_kv", 43120, 3072, 5120),
              ("qkv/cp8", 5390, 15360, 5120), ("o_proj/cp8", 5390, 5120, 5120), ("ffn1/cp8", 5390, 13824, 5120),
              ("ffn2/cp8", 5390, 5120, 13824)]
    if args.quick:
        shapes = shapes[:4] + shapes[5:6]
    g = torch.Generator(device="cuda").manual_seed(1)
    for name, M, N, K in shapes:
        x = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
        b = torch.randn(N, device="cuda", generator=g).to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        flops = 2.0 * M * N * K
        impls = {"dualforce_b200/cg1": lambda: B.ops.linear(x, w, b, out=out, cta_group=1),
                 "dualforce_b200/cg2": lambda: B.ops.linear(x, w, b, out=out, cta_group=2),
                 "F.linear(cuBLASLt)": lambda: F.linear(x, w, b)}
        ref = None
        for impl, fn in impls.items():
            try:
                o = fn()
                torch.cuda.synchronize()
                if ref is None:
                    ref = o.float().clone()
                    err = 0.0
                else:
                    err = (o.float() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
                med, best = time_fn(fn, args.iters, flush)
                emit(op="gemm", name=name, M=M, N=N, K=K, impl=impl, ms=med, ms_best=best, tflops=flops / med * 1e-9,
                     tflops_best=flops / best * 1e-9, max_err_vs_ours=err)
            except Exception as e:  # noqa: BLE001
                emit(op="gemm", name=name, M=M, N=N, K=K, impl=impl, error=repr(e)[:300])
        del x, w, b, out, ref


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", choices=["attention", "gemm"], default=None)
    args = ap.parse_args()
    import dualforce_b200 as B

    B._lib.require_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    emit(op="env", gpu=torch.cuda.get_device_name(0), torch=torch.__version__,
         cudnn=torch.backends.cudnn.version(), iters=args.iters)
    if args.only in (None, "attention"):
        attention_cases(args, B, flush)
    if args.only in (None, "gemm"):
        gemm_cases(args, B, flush)


if __name__ == "__main__":
    main()
