#!/usr/bin/env python
"""Benchmark of the MOVA-360p dual-tower denoising step on B200 (BASELINE.json metric: denoise steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host CPU cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # context parallel (cp_size = N) over NCCL

A "step" is what ``MOVA.__call__`` does per scheduler iteration at cfg_scale > 1 (pipeline_mova.py:416-475): two
``inference_single_step`` calls (positive and negative prompt, B = 1 each: time / text embeddings, patchify, the
dual-tower forward ``forward_dual_tower_dit`` of pipeline_mova.py:612-711, head, unpatchify), classifier-free
guidance and the scheduler's Euler update of both latents -- on the full MOVA-360p geometry: 40 video blocks
(5120 / 40 heads / ffn 13824), 30 audio blocks (1536 / 12 / 8960), 30 bridge layers in both directions,
L_v = 49x22x40 = 43120 video tokens, L_a = 403 audio tokens, 512 text tokens, bf16, random-init weights, synthetic
latents (there is no checkpoint in this sandbox).  Two video experts are resident (``video_dit`` / ``video_dit_2``,
pipeline_mova.py:406-415), as in the reference's memory footprint.

Before anything is timed, every rank runs BASELINE.json configs[0] (2 + 2 blocks at full widths, L_v = 4400) through
the same code path -- context parallel when N > 1 -- and compares it with sampled outputs of the REFERENCE itself
(``dualforce_b200.selfcheck``); the result is the ``parity`` object of the line and a failure exits non-zero.

    python bench.py --gpus N --schedule 50        # BASELINE configs[2]: the whole 50-step denoising loop, timed as one

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
# BASELINE.md section 4 / BASELINE.json configs[0], run AS IS: 2 video + 2 audio blocks + 2 bridge layers at full widths
CPU_CFG = dict(FULL_360P, visual_layers=2, audio_layers=2, grid_size=(5, 22, 40), audio_len=36)
CPU_SAMPLE_TEXT = ("BASELINE configs[0] as is: 2 video + 2 audio blocks + 2 bridge layers (both directions) at full MOVA "
                   "widths (5120/40, 1536/12), 352x640x17-frame clip (L_v=4400, L_a=36, 512 text tokens), fp32, torch "
                   "CPU, all host cores; `measured` is that configuration's own rate, `value` its forward time scaled "
                   "by FLOPs(full 360p CFG step) / FLOPs(configs[0] forward) -- an estimate, labelled as such")


def cpu_sample_runner():
    """Returns (run, scale, cores): run() executes ONE configs[0] forward and returns seconds; ``scale`` x seconds is
    the estimated duration of the full MOVA-360p CFG step (2 forwards) by FLOP count."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch

    import mova_oracle as O  # test infrastructure, used here only as the CPU baseline (kind "port")

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = CPU_CFG
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


def cpu_baseline_object(sec_forward, scale, cores):
    return {"value": 1.0 / (sec_forward * scale), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": CPU_SAMPLE_TEXT,
            "measured": {"config": "BASELINE configs[0] (2+2+2 layers, L_v=4400)", "seconds_per_forward": sec_forward,
                         "steps_per_s": 1.0 / (FORWARDS_PER_STEP * sec_forward), "unit": "configs[0] CFG steps/s"},
            "estimate": {"what": "full MOVA-360p CFG step", "flop_scale": scale,
                         "seconds_per_step": sec_forward * scale}}


def run_reference_arm(args):
    """Each of the K timed 'steps' is one forward of configs[0] (the bounded sample); the line's value is the
    FLOP-scaled estimate for the full step, the measured configs[0] rate is given beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, scale, cores = cpu_sample_runner()
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    times = [run() for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    sec = sum(times) / len(times)
    cb = cpu_baseline_object(sec, scale, cores)
    value = cb["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * scale * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MOVA-360p full dual-tower denoising step (40 video / 30 audio / 30 bridge layers, "
                               "L_v=43120, L_a=403): ESTIMATED by FLOP count from the measured BASELINE configs[0] "
                               "forward (2+2+2 layers, L_v=4400); each timed step of this arm = one configs[0] forward",
                   "measured_seconds_per_configs0_forward": sec, "timed_region_s": wall},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------
def build_model(cfg, device, seed=0, experts=1):
    """WanModel / WanAudioModel / bridge twins with random-init bf16 weights on ``device`` and a pipeline-like
    namespace with both drop-in methods bound.  ``experts=2`` also builds ``video_dit_2`` (the low-noise expert of
    pipeline_mova.py:406-415) so that the memory footprint is the reference's."""
    import types

    import torch

    import dualforce_b200 as B

    torch.manual_seed(seed)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)

    def video_tower():
        return B.WanModel(dim=cfg["visual_dim"], in_dim=STEP_360P["visual_in_dim"], ffn_dim=cfg["visual_ffn"],
                          out_dim=STEP_360P["visual_out_dim"], text_dim=STEP_360P["text_dim"],
                          freq_dim=STEP_360P["freq_dim"], eps=cfg["eps"], patch_size=STEP_360P["visual_patch"],
                          num_heads=cfg["visual_heads"], num_layers=cfg["visual_layers"], has_image_input=False)

    try:
        with torch.device(device):
            vis = video_tower()
            vis2 = video_tower() if experts > 1 else None
            aud = B.WanAudioModel(dim=cfg["audio_dim"], in_dim=STEP_360P["audio_in_dim"], ffn_dim=cfg["audio_ffn"],
                                  out_dim=STEP_360P["audio_out_dim"], text_dim=STEP_360P["text_dim"],
                                  freq_dim=STEP_360P["freq_dim"], eps=cfg["eps"], patch_size=STEP_360P["audio_patch"],
                                  num_heads=cfg["audio_heads"], num_layers=cfg["audio_layers"], has_image_input=False,
                                  vae_type="dac")
            bridge = B.DualTowerConditionalBridge(
                visual_layers=cfg["visual_layers"], audio_layers=cfg["audio_layers"], visual_hidden_dim=cfg["visual_dim"],
                audio_hidden_dim=cfg["audio_dim"], audio_fps=cfg["audio_fps"], head_dim=cfg["head_dim"],
                interaction_strategy=cfg["interaction_strategy"], apply_cross_rope=cfg["apply_cross_rope"])
    finally:
        torch.set_default_dtype(prev)
    pipe = types.SimpleNamespace(video_dit=vis, video_dit_2=vis2, audio_dit=aud, dual_tower_bridge=bridge)
    pipe.forward_dual_tower_dit = types.MethodType(B.forward_dual_tower_dit, pipe)
    pipe.inference_single_step = types.MethodType(B.inference_single_step, pipe)
    return pipe


def host_step_inputs(cfg, seed=2, pin=True):
    """Host tensors of one denoising step as MOVA.__call__ holds them (pipeline_mova.py:378-437): fp32 noise latents
    (16 ch), the mask + first-frame condition (20 ch), fp32 audio latents, the two T5 prompt embeddings
    [1, 512, 4096] bf16."""
    import torch

    g = torch.Generator().manual_seed(seed)
    f, h, w = cfg["grid_size"]
    pt, ph, pw = STEP_360P["visual_patch"]
    n_lat = STEP_360P["visual_out_dim"]
    inp = {"latents": torch.randn(1, n_lat, f * pt, h * ph, w * pw, generator=g),
           "condition": torch.randn(1, STEP_360P["visual_in_dim"] - n_lat, f * pt, h * ph, w * pw, generator=g),
           "audio_latents": torch.randn(1, STEP_360P["audio_in_dim"], cfg["audio_len"], generator=g)}
    for name in ("pos", "neg"):
        c = torch.randn(1, cfg["text_len"], STEP_360P["text_dim"], generator=g)
        c[:, 64:] = 0  # zero-padded T5 tokens (pipeline_mova.py:309-312)
        inp[f"context_{name}"] = c.to(torch.bfloat16)
    return {k: v.pin_memory() for k, v in inp.items()} if pin else inp


def flow_match_schedule(num_steps, shift=5.0, num_train=1000, sigma_min=0.003 / 1.002, sigma_max=1.0):
    """(timesteps, sigmas) of the reference's FlowMatchScheduler.set_timesteps in MOVA's configuration (shift 5,
    extra_one_step; schedulers/flow_match.py:43-62): linspace then sigma' = s*sigma / (1 + (s-1)*sigma).  Only the
    bench needs this (there is no reference scheduler on the GPU box); the product takes the scheduler's own lookups."""
    import torch

    sig = torch.linspace(sigma_max, sigma_min, num_steps + 1)[:-1]
    sig = shift * sig / (1 + (shift - 1) * sig)
    return sig * num_train, sig


def nbytes(t):
    return t.numel() * t.element_size()


def cp_exchange_used(pl) -> str:
    """What the context-parallel runtimes of this process actually used for the Ulysses exchange."""
    used = set()
    for rt in pl._RUNTIMES.values():
        px = getattr(rt, "_px", None)
        used.add("peer windows (CUDA IPC, copy engines + flag words)" if (px and rt.exchange == "peer")
                 else "nccl all_to_all_single")
    return " / ".join(sorted(used)) if used else "none"


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import dualforce_b200 as B
    from dualforce_b200 import _lib, selfcheck
    from dualforce_b200 import step as bstep

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
    from dualforce_b200 import pipeline as _pl

    if args.cp_single_stream:
        _pl.CPRuntime.audio_side_stream = False
    _pl.CPRuntime.exchange = args.cp_exchange
    if args.cp_sets:
        _pl.CPRuntime.set_sizes = tuple(int(v) for v in args.cp_sets.split(","))
    if args.cp_push_streams:
        _pl.CPRuntime.push_streams_n = args.cp_push_streams
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        from torch.distributed.device_mesh import init_device_mesh

        cp_mesh = init_device_mesh("cuda", (world,), mesh_dim_names=("cp",))

    def finish(rc=0):
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            try:
                dist.barrier()
                torch.cuda.synchronize()
                dist.destroy_process_group()
            except Exception:  # noqa: BLE001 -- the line is out; a broken communicator must not change the exit code
                pass
        if rc:
            sys.exit(rc)

    # ---- parity gate: BASELINE configs[0] through this very code path (cp = world) vs the reference's samples ----
    parity = None
    if not args.no_parity:
        parity = selfcheck.reduced_360p_parity(device, cp_mesh)
        if world > 1:  # every rank must agree before anything is timed
            flag = torch.tensor([1 if parity.get("ok") else 0], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            parity["ok_all_ranks"] = bool(flag.item())
            parity["ok"] = parity["ok"] and parity["ok_all_ranks"]
        if not parity["ok"]:
            if rank == 0:
                emit({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "parity": parity,
                      "error": "parity against the reference's configs[0] samples FAILED; nothing was timed"})
            finish(1)
            return

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
    experts = args.experts if args.experts is not None else 2
    pipe = build_model(cfg, device, experts=experts)
    # eager launches by default: graph replay measured within 0.5 % of eager at 1 and 8 GPUs (the GPU, not the host,
    # is the bottleneck), and NCCL communicators captured into a graph can stall process-group teardown
    use_graph = bool(args.cuda_graph)
    pipe.mova_b200_cuda_graph = use_graph

    host = host_step_inputs(cfg)
    n_lat = STEP_360P["visual_out_dim"]
    # the model input cat([latents, condition], dim=1) (pipeline_mova.py:416) is ONE persistent fp32 buffer
    model_input = torch.cat([host["latents"], host["condition"]], dim=1).to(device)
    lat_view = model_input[:, :n_lat]
    audio = host["audio_latents"].to(device)
    lat_next, audio_next = torch.empty_like(lat_view, memory_format=torch.contiguous_format), torch.empty_like(audio)
    ctx = {w_: host[f"context_{w_}"].to(device) for w_ in ("pos", "neg")}  # uploaded once per video (:404-405)
    if args.cfg_merge:
        if world > 1:
            raise SystemExit("--cfg-merge is single-GPU (context parallelism keeps the two-call CFG form)")
        ctx["both"] = torch.cat([ctx["pos"], ctx["neg"]], dim=0)
    n_sched = max(args.schedule or 0, 50)
    timesteps, sigmas = flow_match_schedule(n_sched)
    ts_dev = [timesteps[i].reshape(1).to(device=device, dtype=torch.float32) for i in range(n_sched)]
    ts_host = [timesteps[i].reshape(1).float().pin_memory() for i in range(n_sched)]
    sig = [float(x) for x in sigmas] + [0.0]
    cfg_scale = 5.0
    counter = [0]
    torch.cuda.synchronize()

    def one_step(ts, x_in, a_in, lat_in, lat_out, a_out, idx):
        """2 x inference_single_step + CFG + Euler update of both latents (pipeline_mova.py:416-475)."""
        kw = dict(visual_dit=pipe.video_dit, visual_latents=x_in, audio_latents=a_in, timestep=ts, audio_timestep=None,
                  video_fps=cfg["video_fps"], cp_mesh=cp_mesh)
        if args.cfg_merge:  # the reference's cfg_merge form (pipeline_mova.py:443-445): ONE B = 2 forward per step
            out_v, out_a = pipe.inference_single_step(context=ctx["both"], **kw)
            pos_v, neg_v, pos_a, neg_a = out_v[0:1], out_v[1:2], out_a[0:1], out_a[1:2]
        else:
            pos_v, pos_a = pipe.inference_single_step(context=ctx["pos"], **kw)
            neg_v, neg_a = pipe.inference_single_step(context=ctx["neg"], **kw)
        bstep.guided_update(pos_v, neg_v, lat_in, cfg_scale, sig[idx], sig[idx + 1], out=lat_out)
        bstep.guided_update(pos_a, neg_a, a_in, cfg_scale, sig[idx], sig[idx + 1], out=a_out)

    def step_resident():
        i = counter[0] % n_sched
        counter[0] += 1
        # a new timestep every step, as in the scheduler loop; the updated latents go to a second buffer so that every
        # timed step sees the same inputs
        one_step(ts_dev[i], model_input, audio, lat_view, lat_next, audio_next, i)

    out_host = [torch.empty(lat_next.shape, dtype=torch.float32).pin_memory(),
                torch.empty(audio_next.shape, dtype=torch.float32).pin_memory()]

    def step_e2e():
        i = counter[0] % n_sched
        counter[0] += 1
        lat_view.copy_(host["latents"], non_blocking=True)          # H2D: this step's fp32 latents ...
        audio.copy_(host["audio_latents"], non_blocking=True)
        ts = ts_host[i].to(device, non_blocking=True)               # ... and its timestep
        one_step(ts, model_input, audio, lat_view, lat_next, audio_next, i)
        if rank == 0:  # the result is replicated across the cp ranks: one host copy, on the reporting rank
            out_host[0].copy_(lat_next, non_blocking=True)          # D2H: the updated latents
            out_host[1].copy_(audio_next, non_blocking=True)

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

    if args.schedule:
        run_schedule(args, pipe, cfg, host, ctx, cp_mesh, device, world, rank, timed, parity, full)
        finish(0)
        return

    for _ in range(args.warmup):  # the first one fills the prompt memos (text embedding, per-layer text k / v)
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.LAUNCHES
    if not use_graph:
        _lib._TIMERS = []  # per-launch CUDA events around every attention kernel of the timed region
    ms_total = timed(step_resident, args.steps)
    attn_events, _lib._TIMERS = _lib._TIMERS, None
    launches = _lib.LAUNCHES - launches0
    clocks = sampler.stop() if rank == 0 else None
    roofline_pass = "CUDA events around every video self-attention launch inside the timed region"
    eager_ms = None
    if use_graph:
        # events cannot be recorded inside a replayed graph: time the same launches in one eager step right after
        pipe.mova_b200_cuda_graph = False
        step_resident()
        _lib._TIMERS = []
        eager_ms = timed(step_resident, 1)
        attn_events, _lib._TIMERS = _lib._TIMERS, None
        pipe.mova_b200_cuda_graph = True
        roofline_pass = ("CUDA events around every video self-attention launch of one EAGER step run right after the "
                         "timed (graph-replay) region; eager step %.1f ms" % eager_ms)

    # dominant kernel: video self-attention launches inside the timed region
    f, h, w = cfg["grid_size"]
    L_v = f * h * w
    picked = [ev for ev in (attn_events or []) if ev[3] == L_v and ev[4] == L_v]
    sel = [(e0.elapsed_time(e1), 4.0 * b * hh * sq * skv * dd) for (e0, e1, b, sq, skv, hh, dd) in picked]
    peaks = measured_peaks()
    roofline = None
    if sel:
        # Context parallel with an odd head count per rank runs the attention sets of one layer on two alternating
        # streams, so their launch intervals overlap: the kernel's busy time is the UNION of the intervals (all events
        # share the device clock), not their sum.
        base = picked[0][0]
        spans = sorted((base.elapsed_time(e0), base.elapsed_time(e1)) for (e0, e1, *_rest) in picked)
        busy_ms, cur_a, cur_b = 0.0, spans[0][0], spans[0][1]
        for a, b_ in spans[1:]:
            if a > cur_b:
                busy_ms += cur_b - cur_a
                cur_a, cur_b = a, b_
            else:
                cur_b = max(cur_b, b_)
        busy_ms += cur_b - cur_a
        overlapped = busy_ms < 0.98 * sum(t for t, _ in sel)
        avg_ms = busy_ms / len(sel)
        tf = sum(f_ for _, f_ in sel) / (busy_ms * 1e-3) / 1e12  # launches may differ in heads (cp sets)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "attn_traffic.json")
        if world == 1 and full and os.path.exists(tpath):  # the ncu capture is of the 40-head single-GPU launch
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        roofline = {"bound": "tensor", "kernel": "attn_pair_kernel<2,128,4> (video self-attention, L_v x L_v, %s heads x batch per launch)" % sel_heads(attn_events, L_v),
                    "achieved": tf, "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": tf / peaks["sustained"],
                    "peak_kind": f"{peaks['source']} cuBLAS bf16 sustained (kernel timed inside a long step)",
                    "frac_of_burst_peak": tf / peaks["burst"], "frac_of_nominal_2250": tf / 2250.0,
                    "launches_timed": len(sel), "avg_launch_ms": avg_ms,
                    "share_of_step": busy_ms / (eager_ms if use_graph else ms_total),
                    "launch_intervals_overlap": overlapped,
                    "how": roofline_pass + ("; launches of one layer run on two streams and overlap, so time = union of "
                                            "the launch intervals" if overlapped else ""),
                    "traffic": traffic, "traffic_source": traffic_src}

    # end to end through the public step API with HOST buffers: H2D of the step's latents + timestep, D2H of the result
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    h2d = nbytes(host["latents"]) + nbytes(host["audio_latents"]) + nbytes(ts_host[0])
    d2h = sum(nbytes(t) for t in out_host)

    if rank == 0:
        ms_step = ms_total / args.steps
        value = 1e3 / ms_step
        flops_step = FORWARDS_PER_STEP * flops_forward(cfg)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": ("MOVA-360p full dual-tower DiT denoising step (BASELINE configs[1]; configs[2]'s step "
                                    "when cp_size > 1): 2 x inference_single_step [embeddings, patchify, 40 video "
                                    "blocks 5120/40h/ffn13824 + 30 audio blocks 1536/12h/ffn8960 + 30 bidirectional "
                                    "bridge layers, head, unpatchify] + CFG + Euler update of both latents; L_v=43120 "
                                    "(352x640x193f), L_a=403, 512 text tokens, random init")
                       if full else f"NOT the headline config (debug / BASELINE configs[3]): {cfg}",
                       "cp_size": world, "parallelism": f"cp{world}" if world > 1 else "single GPU",
                       "cp_audio_side_stream": (not args.cp_single_stream) if world > 1 else None,
                       "cp_exchange": cp_exchange_used(_pl) if world > 1 else None,
                       "cp_attention_sets": args.cp_sets or "default",
                       "cp_push_streams": (_pl.CPRuntime.push_streams_n or ("one per peer" if world >= 8 else 1)) if world > 1 else None,
                       "cfg_form": "merged: one B=2 forward per step" if args.cfg_merge else "two B=1 forwards per step",
                       "video_experts_resident": experts,
                       "launch_mode": "cuda graph replay" if use_graph else "eager",
                       "l2_policy": "inputs+weights (~36 GB touched per forward) far exceed the 126 MB L2; no flush needed",
                       "tflop_per_step": flops_step / 1e12},
            "model_tflops": flops_step / (ms_step * 1e-3) / 1e12,
            "model_flops_frac_of_sustained_peak": flops_step / (ms_step * 1e-3) / 1e12 / (peaks["sustained"] * world),
            "clocks": clocks,
            "parity": parity,
            "e2e": {"value": 1e3 / (ms_e2e / args.steps), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "api": "2 x pipe.inference_single_step (pipeline_mova.py:500-609 drop-in) + step.guided_update "
                           "(CFG + FlowMatchPairScheduler.step_from_to): fp32 latents + timestep from pinned host "
                           "memory in, updated fp32 latents out; prompt embeddings resident (uploaded once per video)"},
            "gpu_launches": launches,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            run, scale, cores = cpu_sample_runner()
            sec = run()
            line["cpu_baseline"] = cpu_baseline_object(sec, scale, cores)
        emit(line)
    if use_graph and world > 1:
        # a communicator that was captured into a CUDA graph can block in destroy_process_group(); the line is
        # printed, so leave without tearing NCCL down
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    finish(0)


def run_schedule(args, pipe, cfg, host, ctx, cp_mesh, device, world, rank, timed, parity, full):
    """BASELINE configs[2]: the whole denoising loop of MOVA.__call__ (pipeline_mova.py:405-487) through
    ``dualforce_b200.step.denoising_loop`` -- N scheduler iterations, expert switch at boundary_ratio 0.9 included --
    timed as one region."""
    import torch

    from dualforce_b200 import _lib
    from dualforce_b200 import step as bstep

    n = args.schedule
    timesteps, sigmas = flow_match_schedule(n)
    train_t, train_s = flow_match_schedule(1000)
    pairs = torch.stack([timesteps, timesteps], dim=1)

    def timestep_to_sigma(t):  # FlowMatchPairScheduler.timestep_to_sigma (flow_match_pair.py:208-211)
        return float(train_s[torch.argmin((train_t - float(t)).abs())])

    boundary = 0.9 * 1000

    def pick(t):  # pipeline_mova.py:406-415
        return pipe.video_dit_2 if (pipe.video_dit_2 is not None and t < boundary) else pipe.video_dit

    lat, cond, aud = (host[k].to(device) for k in ("latents", "condition", "audio_latents"))
    res = {}

    def loop(pr):
        res["out"] = bstep.denoising_loop(pipe, lat, cond, aud, ctx["pos"], ctx["neg"], pr, timestep_to_sigma,
                                          cfg["video_fps"], cfg_scale=5.0, cp_mesh=cp_mesh, pick_visual_dit=pick)

    loop(pairs[:2])  # warm-up: memos, NCCL channels
    if pipe.video_dit_2 is not None:
        loop(pairs[-1:])  # and the second expert's packed weights
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    l0 = _lib.LAUNCHES
    ms = timed(lambda: loop(pairs), 1)
    launches = _lib.LAUNCHES - l0
    clocks = sampler.stop() if rank == 0 else None
    finite = bool(torch.isfinite(res["out"][0]).all()) and bool(torch.isfinite(res["out"][1]).all())
    if rank == 0:
        flops_step = FORWARDS_PER_STEP * flops_forward(cfg)
        emit({"metric": METRIC, "value": n / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": n, "warmup": 3,
              "ms_per_step": ms / n, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
              "data": "synthetic",
              "config": {"workload": (f"BASELINE configs[2]: MOVA-360p full {n}-step denoising schedule (shift-5 flow-match "
                                      "sigmas, cfg 5, expert switch at 0.9) through step.denoising_loop, timed as one region")
                         if full else f"NOT the headline config: {cfg}, {n}-step schedule",
                         "cp_size": world, "video_experts_resident": 2 if pipe.video_dit_2 is not None else 1,
                         "tflop_per_step": flops_step / 1e12, "schedule_seconds": ms * 1e-3},
              "model_tflops": flops_step * n / (ms * 1e-3) / 1e12, "clocks": clocks, "parity": parity,
              "outputs_finite": finite, "gpu_launches": launches})


def sel_heads(events, L_v):
    kinds = sorted({b * hh for (_, _, b, sq, skv, hh, dd) in events if sq == L_v and skv == L_v})
    return "/".join(str(k) for k in kinds) if kinds else "0"


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
    ap.add_argument("--cuda-graph", type=int, default=0,
                    help="1: replay the forward as a CUDA graph, 0: eager launches (default)")
    ap.add_argument("--no-parity", action="store_true", help="debug: skip the configs[0] parity gate")
    ap.add_argument("--cp-single-stream", action="store_true",
                    help="A/B: run the replicated audio tower + v2a bridge on the main stream (the round-1 order)")
    ap.add_argument("--cfg-merge", action="store_true",
                    help="A/B: run the CFG pair as ONE batched forward (the reference's cfg_merge form), single GPU")
    ap.add_argument("--cp-exchange", choices=["peer", "nccl"], default="peer",
                    help="Ulysses exchange data path: copy-engine pushes into peer windows + flags (default), or NCCL "
                         "all_to_all_single (round-1 path)")
    ap.add_argument("--cp-push-streams", type=int, default=0,
                    help="A/B: streams the remote pushes of the peer exchange are dealt over (default: the build's)")
    ap.add_argument("--cp-sets", default="", help="A/B: attention set sizes for an odd head count per rank, e.g. 1,3,1")
    ap.add_argument("--experts", type=int, default=None, help="resident video experts (default 2, as in the reference)")
    ap.add_argument("--schedule", type=int, default=0,
                    help="BASELINE configs[2]: time the whole N-step denoising loop (step.denoising_loop) as one region")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
