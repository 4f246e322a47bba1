#!/usr/bin/env python
"""Benchmark of the MOVA-360p dual-tower denoising step on B200 (BASELINE.json metric: denoise steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host CPU cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # context parallel (cp_size = N) over NCCL

A "step" is what ``MOVA.__call__`` does per scheduler iteration at cfg_scale > 1 (pipeline_mova.py:429-456): two
forwards of the dual-tower DiT (positive and negative prompt, B = 1 each) through ``forward_dual_tower_dit``
(pipeline_mova.py:612-711) on the full MOVA-360p geometry: 40 video blocks (5120 / 40 heads / ffn 13824), 30 audio
blocks (1536 / 12 / 8960), 30 bridge layers in both directions, L_v = 49x22x40 = 43120 video tokens, L_a = 403 audio
tokens, 512 text tokens, bf16, random-init weights, synthetic latents (there is no checkpoint in this sandbox).

One JSON line is printed by rank 0; see DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FULL_360P = dict(visual_dim=5120, visual_heads=40, visual_ffn=13824, visual_layers=40, audio_dim=1536, audio_heads=12,
                 audio_ffn=8960, audio_layers=30, head_dim=128, interaction_strategy="full", apply_cross_rope=True,
                 audio_fps=50.0, grid_size=(49, 22, 40), audio_len=403, text_len=512, eps=1e-6, video_fps=24.0)
# tower-level hyper-parameters of the MOVA-360p checkpoint around the blocks
STEP_360P = dict(visual_in_dim=36, visual_out_dim=16, visual_patch=(1, 2, 2), audio_in_dim=128, audio_out_dim=128,
                 audio_patch=(1,), text_dim=4096, freq_dim=256)
METRIC = "denoise steps/sec, MOVA-360p dual-tower DiT (2 CFG forwards per step)"
UNIT = "steps/s"
FORWARDS_PER_STEP = 2


def flops_forward(cfg) -> float:
    """Algorithmic FLOPs of one forward: GEMM 2MNK, attention 4*H*Sq*Skv*D."""
    f, h, w = cfg["grid_size"]
    L_v, L_a, L_t = f * h * w, cfg["audio_len"], cfg["text_len"]
    dv, da = cfg["visual_dim"], cfg["audio_dim"]

    def block(L, d, ffn):
        return 2 * L * d * d * 6 + 2 * L_t * d * d * 2 + 2 * L * d * ffn * 2 + 4 * L * L * d + 4 * L * L_t * d

    n_inter = min(cfg["visual_layers"], cfg["audio_layers"])
    a2v = 2 * L_v * dv * dv * 2 + 2 * L_a * da * dv * 2 + 4 * L_v * L_a * dv
    v2a = 2 * L_a * da * da * 2 + 2 * L_v * dv * da * 2 + 4 * L_a * L_v * da
    return (cfg["visual_layers"] * block(L_v, dv, cfg["visual_ffn"]) + cfg["audio_layers"] * block(L_a, da, cfg["audio_ffn"])
            + n_inter * (a2v + v2a))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), hbm=p["hbm_gbs"],
                    source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 200 ms while the timed region runs (nvidia-smi)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------------------------
CPU_SAMPLE_CFG = dict(FULL_360P, visual_layers=1, audio_layers=1, grid_size=(5, 22, 40), audio_len=36)
CPU_SAMPLE_TEXT = ("1 video block + 1 audio block + 1 bridge layer (both directions) at full MOVA widths "
                   "(5120/40, 1536/12), 352x640x17-frame clip (L_v=4400, L_a=36, 512 text tokens), fp32, "
                   "torch CPU; steps/s = 1 / (sample seconds x FLOPs(full 360p step) / FLOPs(sample))")


def cpu_sample_runner():
    """Returns (run, scale): run() executes the bounded sample once and returns seconds; scale converts to the
    full-step estimate."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch

    import mova_oracle as O  # test infrastructure, used here only as the CPU baseline (kind "port")

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = CPU_SAMPLE_CFG
    Pv, Pa, Pb, inp = O.make_case(cfg, 0)

    def run():
        t0 = time.perf_counter()
        with torch.no_grad():
            O.forward_dual_tower_dit(Pv, Pa, Pb, cfg, inp["visual_x"], inp["audio_x"], inp["visual_context"],
                                     inp["audio_context"], inp["visual_t_mod"], inp["audio_t_mod"], inp["visual_freqs"],
                                     inp["audio_freqs"], cfg["grid_size"], cfg["video_fps"])
        return time.perf_counter() - t0

    scale = FORWARDS_PER_STEP * flops_forward(FULL_360P) / flops_forward(cfg)
    return run, scale, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, scale, cores = cpu_sample_runner()
    for _ in range(args.warmup):
        run()
    times = [run() for _ in range(args.steps)]
    sec = sum(times) / len(times)
    value = 1.0 / (sec * scale)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * scale * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MOVA-360p full dual-tower denoising step (40 video / 30 audio / 30 bridge layers, "
                               "L_v=43120, L_a=403), estimated from a bounded CPU sample", "sample_seconds": sec},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": CPU_SAMPLE_TEXT},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------
def build_model(cfg, device, seed=0, with_step=False):
    import torch

    import dualforce_b200 as B

    torch.manual_seed(seed)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            vis = torch.nn.Module()
            vis.blocks = torch.nn.ModuleList([B.DiTBlock(False, cfg["visual_dim"], cfg["visual_heads"], cfg["visual_ffn"],
                                                         cfg["eps"]) for _ in range(cfg["visual_layers"])])
            aud = torch.nn.Module()
            aud.blocks = torch.nn.ModuleList([B.DiTBlock(False, cfg["audio_dim"], cfg["audio_heads"], cfg["audio_ffn"],
                                                         cfg["eps"]) for _ in range(cfg["audio_layers"])])
            if with_step:
                # the rest of WanModel / WanAudioModel: embeddings, patch embedding, head -- what
                # MOVA.inference_single_step drives around the path.  Zero-layer twins donate those parameters; the
                # block lists above stay exactly as the forward-level measurement has always built them.
                tv = B.WanModel(dim=cfg["visual_dim"], in_dim=STEP_360P["visual_in_dim"], ffn_dim=cfg["visual_ffn"],
                                out_dim=STEP_360P["visual_out_dim"], text_dim=STEP_360P["text_dim"],
                                freq_dim=STEP_360P["freq_dim"], eps=cfg["eps"], patch_size=STEP_360P["visual_patch"],
                                num_heads=cfg["visual_heads"], num_layers=0, has_image_input=False)
                ta = B.WanAudioModel(dim=cfg["audio_dim"], in_dim=STEP_360P["audio_in_dim"], ffn_dim=cfg["audio_ffn"],
                                     out_dim=STEP_360P["audio_out_dim"], text_dim=STEP_360P["text_dim"],
                                     freq_dim=STEP_360P["freq_dim"], eps=cfg["eps"], patch_size=STEP_360P["audio_patch"],
                                     num_heads=cfg["audio_heads"], num_layers=0, has_image_input=False, vae_type="dac")
                tv.blocks, ta.blocks = vis.blocks, aud.blocks
                vis, aud = tv, ta
            bridge = B.DualTowerConditionalBridge(
                visual_layers=cfg["visual_layers"], audio_layers=cfg["audio_layers"], visual_hidden_dim=cfg["visual_dim"],
                audio_hidden_dim=cfg["audio_dim"], audio_fps=cfg["audio_fps"], head_dim=cfg["head_dim"],
                interaction_strategy=cfg["interaction_strategy"], apply_cross_rope=cfg["apply_cross_rope"])
    finally:
        torch.set_default_dtype(prev)
    import types

    pipe = types.SimpleNamespace(video_dit=vis, video_dit_2=None, audio_dit=aud, dual_tower_bridge=bridge)
    pipe.forward_dual_tower_dit = types.MethodType(B.forward_dual_tower_dit, pipe)
    if with_step:
        pipe.inference_single_step = types.MethodType(B.inference_single_step, pipe)
    return pipe


def host_step_inputs(cfg, seed=2, pin=True):
    """Pinned host tensors of one denoising step as MOVA.__call__ hands them to inference_single_step
    (pipeline_mova.py:416-437): fp32 latents (16 noise + 20 condition channels), fp32 audio latents, the two T5
    prompt embeddings [1, 512, 4096] bf16, the timestep."""
    import torch

    g = torch.Generator().manual_seed(seed)
    f, h, w = cfg["grid_size"]
    pt, ph, pw = STEP_360P["visual_patch"]
    inp = {"visual_latents": torch.randn(1, STEP_360P["visual_in_dim"], f * pt, h * ph, w * pw, generator=g),
           "audio_latents": torch.randn(1, STEP_360P["audio_in_dim"], cfg["audio_len"], generator=g),
           "timestep": torch.tensor([900.0], dtype=torch.float32)}
    for name in ("pos", "neg"):
        c = torch.randn(1, cfg["text_len"], STEP_360P["text_dim"], generator=g)
        c[:, 64:] = 0  # zero-padded T5 tokens (pipeline_mova.py:309-312)
        inp[f"context_{name}"] = c.to(torch.bfloat16)
    return {k: v.pin_memory() for k, v in inp.items()} if pin else inp


def host_inputs(cfg, seed=1):
    """Pinned host tensors of one step: shared latents/tables plus a positive and a negative text context."""
    import torch

    import dualforce_b200 as B

    g = torch.Generator().manual_seed(seed)
    f, h, w = cfg["grid_size"]
    L_v, L_a, L_t = f * h * w, cfg["audio_len"], cfg["text_len"]
    dv, da = cfg["visual_dim"], cfg["audio_dim"]

    def rn(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).pin_memory()

    inp = {"visual_x": rn(1, L_v, dv), "audio_x": rn(1, L_a, da), "visual_t_mod": rn(1, 6, dv, scale=0.3),
           "audio_t_mod": rn(1, 6, da, scale=0.3)}
    for name in ("pos", "neg"):
        cv, ca = torch.randn(1, L_t, dv, generator=g), torch.randn(1, L_t, da, generator=g)
        cv[:, 64:] = 0  # zero-padded T5 tokens (pipeline_mova.py:309-312)
        ca[:, 64:] = 0
        inp[f"visual_context_{name}"] = cv.to(torch.bfloat16).pin_memory()
        inp[f"audio_context_{name}"] = ca.to(torch.bfloat16).pin_memory()
    inp["visual_freqs"] = B.rope.video_freqs(B.rope.precompute_freqs_cis_3d(cfg["head_dim"]), cfg["grid_size"], "cpu").contiguous().pin_memory()
    inp["audio_freqs"] = B.rope.audio_freqs(B.rope.precompute_freqs_cis_1d(cfg["head_dim"]), L_a, "cpu").contiguous().pin_memory()
    return inp


def nbytes(t):
    return t.numel() * t.element_size()


def measure_step_api(pipe, cfg, device, cp_mesh, rank, steps, timed, launches, pin=True):
    """The e2e leg through the step-level public API, the call MOVA.__call__ makes per scheduler iteration
    (pipeline_mova.py:429-456): 2 x ``pipe.inference_single_step`` with fp32 latents + a fresh timestep copied from
    pinned host memory inside the timed region and the bf16 denoised latents copied back (reporting rank only; the
    result is replicated across cp ranks).  ``timed(fn, steps)`` is the bench's barrier / CUDA-event timer,
    ``launches()`` the running count of kernel launches.  Returns the ``e2e`` object of the JSON line."""
    import torch

    def pinned(t):
        return t.pin_memory() if pin else t

    host_s = {k: pinned(v) for k, v in host_step_inputs(cfg, pin=False).items()}
    out_host = []
    # a new timestep every step, like the scheduler loop (one pinned scalar each: an async copy never reads a buffer
    # the host has since rewritten)
    ts_host = [pinned(torch.tensor([900.0 - 18.0 * i], dtype=torch.float32)) for i in range(steps + 4)]
    ts_next = [0]
    # the two prompt embeddings are uploaded once per video, as in MOVA.__call__ (:404-405), not once per step
    ctx_dev = {w_: host_s[f"context_{w_}"].to(device) for w_ in ("pos", "neg")}

    def step_e2e_api():
        ts = ts_host[ts_next[0] % len(ts_host)]
        ts_next[0] += 1
        d = {k: host_s[k].to(device, non_blocking=True) for k in ("visual_latents", "audio_latents")}
        d["timestep"] = ts.to(device, non_blocking=True)
        outs = []
        for which in ("pos", "neg"):
            outs.append(pipe.inference_single_step(
                visual_dit=pipe.video_dit, visual_latents=d["visual_latents"], audio_latents=d["audio_latents"],
                context=ctx_dev[which], timestep=d["timestep"], audio_timestep=None, video_fps=cfg["video_fps"],
                cp_mesh=cp_mesh))
        if not out_host:
            out_host.extend([pinned(torch.empty(t.shape, dtype=t.dtype)) for t in pair] for pair in outs)
        if rank == 0:
            for pair, hpair in zip(outs, out_host):
                for t, ht in zip(pair, hpair):
                    ht.copy_(t, non_blocking=True)
        return outs

    step_e2e_api()  # first step: fills the prompt memos (text embedding, per-layer text k / v)
    step_e2e_api()
    before = launches()
    step_e2e_api()
    launches_per_step = launches() - before  # one steady-state step
    ms_api = timed(step_e2e_api, steps)
    per_step_in = sum(nbytes(host_s[k]) for k in ("visual_latents", "audio_latents")) + nbytes(ts_host[0])
    return {"value": 1e3 / (ms_api / steps), "unit": UNIT, "h2d_bytes_per_step": per_step_in,
            "d2h_bytes_per_step": sum(nbytes(t) for pair in out_host for t in pair), "ms_per_step": ms_api / steps,
            "gpu_launches_per_step": launches_per_step,
            "api": "2 x pipe.inference_single_step (pipeline_mova.py:500-609 drop-in): fp32 latents + timestep from "
                   "pinned host memory in, bf16 denoised latents out; prompt embeddings resident (uploaded once per video)"}


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import dualforce_b200 as B
    from dualforce_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (one process per GPU)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    _lib.require_device(local_rank)
    cp_mesh = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        from torch.distributed.device_mesh import init_device_mesh

        cp_mesh = init_device_mesh("cuda", (world,), mesh_dim_names=("cp",))

    cfg = dict(FULL_360P)
    if args.video_layers is not None:
        cfg["visual_layers"] = args.video_layers
    if args.audio_layers is not None:
        cfg["audio_layers"] = args.audio_layers
    if args.frames is not None:
        cfg["grid_size"] = ((args.frames - 1) // 4 + 1, 22, 40)
    if args.res == "720p":  # BASELINE.json configs[3]: 720x1280 -> token grid (f, 45, 80), L_v = 176400 at 193 frames
        cfg["grid_size"] = (cfg["grid_size"][0], 45, 80)
    full = (cfg == FULL_360P)

    step_api_error = None
    want_step_api = os.environ.get("MOVA_BENCH_STEP_API", "1") != "0"
    try:
        pipe = build_model(cfg, device, with_step=want_step_api)
    except Exception as exc:  # the forward-level measurement must not depend on the step wrapper
        if not want_step_api:
            raise
        step_api_error = f"build: {exc!r}"[:300]
        pipe = build_model(cfg, device, with_step=False)
    # eager launches by default: graph replay measured within 0.5 % of eager at 1 and 8 GPUs (the GPU, not the host,
    # is the bottleneck), and NCCL communicators captured into a graph can stall process-group teardown
    use_graph = bool(args.cuda_graph) if args.cuda_graph is not None else False
    pipe.mova_b200_cuda_graph = use_graph
    host = host_inputs(cfg)
    dev = {k: v.to(device, non_blocking=True) for k, v in host.items()}
    torch.cuda.synchronize()

    def forward(d, which):
        return pipe.forward_dual_tower_dit(
            pipe.video_dit, d["visual_x"], d["audio_x"], d[f"visual_context_{which}"], d[f"audio_context_{which}"],
            d["visual_t_mod"], d["audio_t_mod"], d["visual_freqs"], d["audio_freqs"], cfg["grid_size"], cfg["video_fps"],
            cp_mesh=cp_mesh)

    def step_resident():
        outs = []
        for which in ("pos", "neg"):
            outs.append(forward(dev, which))
        return outs

    out_host = None

    def step_e2e():
        nonlocal out_host
        d = {k: v.to(device, non_blocking=True) for k, v in host.items()}
        outs = []
        for which in ("pos", "neg"):
            outs.append(forward(d, which))
        if out_host is None:
            out_host = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in pair] for pair in outs]
        if rank == 0:  # the result is replicated across the cp ranks: one host copy, on the reporting rank
            for pair, hpair in zip(outs, out_host):
                for t, ht in zip(pair, hpair):
                    ht.copy_(t, non_blocking=True)
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """EXACTLY `steps` steps bracketed by barrier + synchronize; device time (CUDA events), max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    launches0 = _lib.LAUNCHES
    step_resident()  # first warm-up step: in graph mode this one captures the graph
    launches_per_step = _lib.LAUNCHES - launches0
    if use_graph:  # capture = 2 eager warm-ups + 1 recorded pass of ONE forward; a step replays it twice
        launches_per_step = (launches_per_step // 3) * FORWARDS_PER_STEP
    for _ in range(args.warmup - 1):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if not use_graph:
        _lib._TIMERS = []  # per-launch CUDA events around every attention kernel of the timed region
    ms_total = timed(step_resident, args.steps)
    launches = launches_per_step * args.steps
    attn_events, _lib._TIMERS = _lib._TIMERS, None
    clocks = sampler.stop() if rank == 0 else None
    roofline_pass = "CUDA events around every video self-attention launch inside the timed region"
    if use_graph:
        # events cannot be recorded inside a replayed graph: time the same launches in one eager step right after
        pipe.mova_b200_cuda_graph = False
        step_resident()
        _lib._TIMERS = []
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_resident()
        e1.record()
        barrier()
        attn_events, _lib._TIMERS = _lib._TIMERS, None
        eager_ms = e0.elapsed_time(e1)
        pipe.mova_b200_cuda_graph = True
        roofline_pass = ("CUDA events around every video self-attention launch of one EAGER step run right after the "
                         "timed (graph-replay) region; eager step %.1f ms" % eager_ms)

    # dominant kernel: video self-attention launches inside the timed region
    heads_local = cfg["visual_heads"] // world if world > 1 else cfg["visual_heads"]
    f, h, w = cfg["grid_size"]
    L_v = f * h * w
    sel = [(e0.elapsed_time(e1), 4.0 * b * hh * sq * skv * dd) for (e0, e1, b, sq, skv, hh, dd) in attn_events
           if sq == L_v and skv == L_v]
    del heads_local
    peaks = measured_peaks()
    roofline = None
    if sel:
        avg_ms = sum(t for t, _ in sel) / len(sel)
        tf = sel[0][1] / (avg_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "attn_traffic.json")
        if world == 1 and os.path.exists(tpath):  # the ncu capture is of the 40-head single-GPU launch
            with open(tpath) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": "attn_fwd_kernel (video self-attention, L_v x L_v, %d heads/launch)" % sel_heads(attn_events, L_v),
                    "achieved": tf, "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": tf / peaks["sustained"],
                    "peak_kind": f"{peaks['source']} cuBLAS bf16 sustained (kernel timed inside a long step)",
                    "frac_of_burst_peak": tf / peaks["burst"], "frac_of_nominal_2250": tf / 2250.0,
                    "launches_timed": len(sel), "avg_launch_ms": avg_ms,
                    "share_of_step": sum(t for t, _ in sel) / (eager_ms if use_graph else ms_total),
                    "how": roofline_pass,
                    "traffic": traffic}

    # end to end through the public API with host buffers (H2D of the step's inputs, D2H of its outputs)
    for _ in range(1):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    h2d = sum(nbytes(v) for v in host.values())
    d2h = sum(nbytes(t) for pair in out_host for t in pair)

    # ... and through the step-level public API, the call MOVA.__call__ makes (pipeline_mova.py:429-456):
    # latents / prompt embeddings / timestep from pinned host memory in, denoised latents out.  Measured last and
    # guarded: a failure here is reported in the line and leaves every number above untouched.
    e2e_step = None
    if want_step_api and step_api_error is None:
        try:
            e2e_step = measure_step_api(pipe, cfg, device, cp_mesh, rank, args.steps, timed, lambda: _lib.LAUNCHES)
        except Exception as exc:
            step_api_error = repr(exc)[:300]

    if rank == 0:
        ms_step = ms_total / args.steps
        value = 1e3 / ms_step
        flops_step = FORWARDS_PER_STEP * flops_forward(cfg)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": ("MOVA-360p full dual-tower DiT denoising step: 2 CFG forwards x (40 video blocks "
                                    "5120/40h/ffn13824 + 30 audio blocks 1536/12h/ffn8960 + 30 bidirectional bridge "
                                    "layers), L_v=43120 (352x640x193f), L_a=403, 512 text tokens, random init")
                       if full else f"NOT the headline config (debug / BASELINE configs[3]): {cfg}",
                       "cp_size": world, "parallelism": f"cp{world}" if world > 1 else "single GPU",
                       "launch_mode": "cuda graph replay" if use_graph else "eager",
                       "l2_policy": "inputs+weights (~36 GB touched per forward) far exceed the 126 MB L2; no flush needed",
                       "tflop_per_step": flops_step / 1e12},
            "model_tflops": flops_step / (ms_step * 1e-3) / 1e12,
            "model_flops_frac_of_sustained_peak": flops_step / (ms_step * 1e-3) / 1e12 / (peaks["sustained"] * world),
            "clocks": clocks,
            "e2e": {"value": 1e3 / (ms_e2e / args.steps), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "api": "2 x pipe.forward_dual_tower_dit (pipeline_mova.py:612-711 drop-in): token-level hidden "
                           "states, contexts, t_mod and RoPE tables from pinned host memory in, hidden states out"},
            "gpu_launches": launches,
            "roofline": roofline,
        }
        if e2e_step is not None:
            # the headline end-to-end number is the call a user of the reference makes per scheduler iteration
            line["e2e_forward_api"] = line["e2e"]
            line["e2e"] = e2e_step
        if step_api_error is not None:
            line["e2e_step_api_error"] = step_api_error
        if world == 1 and not args.no_cpu_baseline:
            run, scale, cores = cpu_sample_runner()
            run()  # warm-up (thread pool, allocator)
            sec = run()
            line["cpu_baseline"] = {"value": 1.0 / (sec * scale), "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": CPU_SAMPLE_TEXT, "sample_seconds": sec}
        emit(line)
    if world == 1 and step_api_error is not None:  # a failed launch leaves a sticky context error: skip teardown
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    if world > 1:
        if step_api_error is not None:  # the device / communicator may be unusable: the line is out, leave
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.barrier()
        torch.cuda.synchronize()
        if use_graph:
            # a communicator that was captured into a CUDA graph can block in destroy_process_group(); the line is
            # printed and every rank has passed the barrier, so leave without tearing NCCL down
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


def sel_heads(events, L_v):
    for (_, _, b, sq, skv, hh, dd) in events:
        if sq == L_v and skv == L_v:
            return hh
    return 0


_JSON_FD = None


def claim_stdout():
    """Keep stdout clean for the ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    chatter) is sent to stderr; emit() writes the result to the original stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj) -> None:
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--video-layers", type=int, default=None, help="debug: reduced depth (marks the line REDUCED)")
    ap.add_argument("--audio-layers", type=int, default=None)
    ap.add_argument("--frames", type=int, default=None, help="debug: clip length in frames (default 193)")
    ap.add_argument("--res", choices=["360p", "720p"], default="360p",
                    help="720p = BASELINE.json configs[3] (L_v = 176400; meant for --gpus 8); marks the line REDUCED/other")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-graph", type=int, default=None,
                    help="1: replay the forward as a CUDA graph, 0: eager launches (default)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
