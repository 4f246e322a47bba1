"""The denoising step around the dual-tower forward: drop-in for ``MOVA.inference_single_step``
(mova/diffusion/pipelines/pipeline_mova.py:500-609) plus B200 twins of the tower containers it drives --
``WanModel`` (mova/diffusion/models/wan_video_dit.py:333-416), ``WanAudioModel``
(mova/diffusion/models/wan_audio_dit.py:108-195) and ``Head`` (wan_video_dit.py:314-330).

What the reference does per call, and what happens here instead:

=====================================  =========================================================================
reference (per call, 100x per video)    here
=====================================  =========================================================================
time_embedding / time_projection        ``ops.sinusoidal_embedding`` + three fp32 GEMV launches (the reference runs
under autocast(float32) (:544-549)      them in fp32 too); memoised per timestep tensor, so the negative-prompt call
                                        of the same step reuses the positive call's result
text_embedding MLP (:558-559)           two tcgen05 GEMMs (GELU-tanh fused); memoised per context tensor -- the two
                                        prompts are constant over the 50 steps
per-layer text k / v (:218-223)         memoised inside ``CrossAttention.kv`` for contexts produced by this file
Conv3d / Conv1d patchify (:561, :573)   ``ops.patchify`` (im2col + bf16 cast, one pass) + one tcgen05 GEMM
RoPE table assembly (:566-585)          memoised per (grid, device)
head (LN + modulate + Linear) (:603)    ``ops.layernorm`` (modulate fused) + tcgen05 GEMM, run on the LOCAL token
                                        chunk under context parallelism
all-gather of [L_v, 5120] (:704-706)    all-gather of the head output [L_v, 64] (80x fewer bytes)
unpatchify (:604, :607)                 ``ops.unpatchify``
=====================================  =========================================================================

There is no PyTorch compute path: every tensor op above is a kernel of ``libmova_b200.so``; torch supplies memory,
views, the NCCL all-gather and the caches' dictionaries.
"""
from __future__ import annotations

import math
import types
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import cp as cpmod
from . import ops, rope
from .modules import DiTBlock, param_sig

__all__ = ["Head", "WanModel", "WanAudioModel", "inference_single_step", "embed_time", "embed_text", "patchify",
           "head_unpatchify", "guided_update", "denoising_loop", "clear_step_caches"]

_STATIC_FLAG = "_mova_b200_static"  # set on context embeddings whose per-layer k / v may be memoised


def _version(t: torch.Tensor) -> int:
    try:
        return t._version
    except RuntimeError:  # inference-mode tensors carry no version counter (and cannot be modified in place)
        return -1


def _tensor_key(t: torch.Tensor) -> tuple:
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, _version(t), str(t.device))


class _Memo:
    """Tiny LRU keyed on tensor identity (pointer, shape, version).  The key tensors are kept alive by the entry, so
    their storage cannot be recycled for different contents while the entry exists."""

    def __init__(self, capacity: int):
        self.capacity = capacity
        self.entries = {}

    def get(self, key):
        hit = self.entries.get(key)
        if hit is None:
            return None
        self.entries[key] = self.entries.pop(key)  # most recently used last
        return hit[1]

    def put(self, key, keep_alive, value):
        while len(self.entries) >= self.capacity:
            self.entries.pop(next(iter(self.entries)))
        self.entries[key] = (keep_alive, value)
        return value

    def clear(self):
        self.entries.clear()


def _memo(owner, name: str, capacity: int) -> _Memo:
    m = owner.__dict__.get(name)
    if m is None:
        m = owner.__dict__[name] = _Memo(capacity)
    return m


def clear_step_caches(*models) -> None:
    """Drop every per-model memo (time / text embeddings, RoPE tables, per-layer text k / v)."""
    for model in models:
        if model is None:
            continue
        for name in ("_mova_b200_time", "_mova_b200_text", "_mova_b200_freqs"):
            m = model.__dict__.get(name)
            if m is not None:
                m.clear()
        for blk in getattr(model, "blocks", []):
            cache = getattr(getattr(blk, "cross_attn", None), "_kv_memo", None)
            if cache is not None:
                cache.clear()


# ----------------------------------------------------------------------------------------------------------------
# module twins
# ----------------------------------------------------------------------------------------------------------------
class Head(nn.Module):
    """wan_video_dit.py:314-330 / wan_audio_dit.py:83-102 -- ``forward(x, t_mod)`` with ``t_mod`` the ``[B, dim]``
    time embedding: ``Linear(LayerNorm(x) * (1 + scale) + shift)``, (shift, scale) = modulation + t."""

    def __init__(self, dim: int, out_dim: int, patch_size: Sequence[int], eps: float):
        super().__init__()
        self.dim = dim
        self.patch_size = patch_size
        self.norm = nn.LayerNorm(dim, eps=eps, elementwise_affine=False)
        self.head = nn.Linear(dim, out_dim * math.prod(patch_size))
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)

    @classmethod
    def from_reference(cls, ref: nn.Module) -> "Head":
        """Wrap a reference ``Head`` sharing its Parameter objects."""
        h = cls.__new__(cls)
        nn.Module.__init__(h)
        h.dim, h.patch_size = ref.dim, ref.patch_size
        h.norm, h.head, h.modulation = ref.norm, ref.head, ref.modulation
        return h

    def forward(self, x: torch.Tensor, t_mod: torch.Tensor) -> torch.Tensor:
        if t_mod.dim() == 3:
            raise NotImplementedError("per-token time embedding (seperated_timestep, wan_video_dit.py:324-326) is not "
                                      "used by MOVA")
        if t_mod.shape[0] != 1:
            raise NotImplementedError("Head: one time embedding per call (a CFG pair shares its timestep)")
        mod = ops.add_to_f32(self.modulation.to(dtype=torch.bfloat16), t_mod.to(torch.bfloat16).unsqueeze(1))
        mod = mod.reshape(2, self.dim)
        h = ops.layernorm(x, self.norm.eps, shift=mod[0], scale=mod[1])
        return ops.linear(h, self.head.weight, self.head.bias)


def _text_embedding(dim: int, text_dim: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(text_dim, dim), nn.GELU(approximate="tanh"), nn.Linear(dim, dim))


def _time_embedding(dim: int, freq_dim: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))


class _Tower(nn.Module):
    """What WanModel and WanAudioModel share (constructor arguments as in the reference; the options MOVA's
    checkpoints leave off raise)."""

    _repeated_blocks = ("DiTBlock",)

    def _init_common(self, dim, in_dim, ffn_dim, out_dim, text_dim, freq_dim, eps, patch_size, num_heads, num_layers,
                     has_image_input, has_image_pos_emb, has_ref_conv, add_control_adapter, seperated_timestep,
                     require_vae_embedding, require_clip_embedding, fuse_vae_embedding_in_latents):
        if has_image_input or has_ref_conv or add_control_adapter:
            raise NotImplementedError("has_image_input / has_ref_conv / add_control_adapter are off in the MOVA "
                                      "checkpoints (SURVEY 0.1) and not built here")
        self.dim, self.in_dim, self.out_dim = dim, in_dim, out_dim
        self.freq_dim = freq_dim
        self.has_image_input = has_image_input
        self.patch_size = tuple(patch_size)
        self.seperated_timestep = seperated_timestep
        self.require_vae_embedding = require_vae_embedding
        self.require_clip_embedding = require_clip_embedding
        self.fuse_vae_embedding_in_latents = fuse_vae_embedding_in_latents
        self.text_embedding = _text_embedding(dim, text_dim)
        self.time_embedding = _time_embedding(dim, freq_dim)
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, dim * 6))
        self.blocks = nn.ModuleList([DiTBlock(has_image_input, dim, num_heads, ffn_dim, eps) for _ in range(num_layers)])
        self.head = Head(dim, out_dim, patch_size, eps)
        self.has_image_pos_emb = has_image_pos_emb
        self.has_ref_conv = has_ref_conv
        self.control_adapter = None

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def patchify(self, x: torch.Tensor, control_camera_latents_input=None):
        if control_camera_latents_input is not None:
            raise NotImplementedError("control adapter (wan_video_dit.py:403-406) is not part of MOVA")
        return patchify(self, x)

    def unpatchify(self, x: torch.Tensor, grid_size):
        return torch.stack([ops.unpatchify(x[b], grid_size, self.patch_size, self.out_dim) for b in range(x.shape[0])])

    @torch.no_grad()
    def forward(self, x: torch.Tensor, timestep: torch.Tensor, context: torch.Tensor, **kwargs) -> torch.Tensor:
        """Single-tower forward (wan_video_dit.py:418-473 / wan_audio_dit.py:197-252) on the same kernels."""
        t, t_mod = embed_time(self, timestep)
        ctx = embed_text(self, context)
        tokens, grid = patchify(self, x)
        freqs = token_freqs(self, grid, tokens.device)
        for block in self.blocks:
            tokens = block(tokens, ctx, t_mod, freqs)
        return head_unpatchify(self, tokens, t, grid)


class WanModel(_Tower):
    """wan_video_dit.py:333-416: Conv3d patch embedding, 3-D RoPE tables, DiTBlocks, Head."""

    def __init__(self, dim: int, in_dim: int, ffn_dim: int, out_dim: int, text_dim: int, freq_dim: int, eps: float,
                 patch_size: Tuple[int, int, int], num_heads: int, num_layers: int, has_image_input: bool,
                 has_image_pos_emb: bool = False, has_ref_conv: bool = False, add_control_adapter: bool = False,
                 in_dim_control_adapter: int = 24, seperated_timestep: bool = False,
                 require_vae_embedding: bool = True, require_clip_embedding: bool = True,
                 fuse_vae_embedding_in_latents: bool = False):
        super().__init__()
        self._init_common(dim, in_dim, ffn_dim, out_dim, text_dim, freq_dim, eps, patch_size, num_heads, num_layers,
                          has_image_input, has_image_pos_emb, has_ref_conv, add_control_adapter, seperated_timestep,
                          require_vae_embedding, require_clip_embedding, fuse_vae_embedding_in_latents)
        self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.freqs = rope.precompute_freqs_cis_3d(dim // num_heads)


class WanAudioModel(_Tower):
    """wan_audio_dit.py:108-195: Conv1d patch embedding, 1-D RoPE tables (``vae_type`` picks the table family)."""

    def __init__(self, dim: int, in_dim: int, ffn_dim: int, out_dim: int, text_dim: int, freq_dim: int, eps: float,
                 patch_size: Sequence[int], num_heads: int, num_layers: int, has_image_input: bool,
                 has_image_pos_emb: bool = False, has_ref_conv: bool = False, add_control_adapter: bool = False,
                 in_dim_control_adapter: int = 24, seperated_timestep: bool = False,
                 require_vae_embedding: bool = True, require_clip_embedding: bool = True,
                 fuse_vae_embedding_in_latents: bool = False, vae_type: str = "oobleck"):
        super().__init__()
        self._init_common(dim, in_dim, ffn_dim, out_dim, text_dim, freq_dim, eps, patch_size, num_heads, num_layers,
                          has_image_input, has_image_pos_emb, has_ref_conv, add_control_adapter, seperated_timestep,
                          require_vae_embedding, require_clip_embedding, fuse_vae_embedding_in_latents)
        self.vae_type = vae_type
        self.patch_embedding = nn.Conv1d(in_dim, dim, kernel_size=self.patch_size, stride=self.patch_size)
        head_dim = dim // num_heads
        if vae_type == "oobleck":
            self.freqs = rope.legacy_precompute_freqs_cis_1d(head_dim, base_tps=4.0, target_tps=44100 / 2048)
        elif vae_type == "dac":
            self.freqs = rope.precompute_freqs_cis_1d(head_dim)
        else:
            raise ValueError(f"Invalid VAE type: {vae_type}")


# ----------------------------------------------------------------------------------------------------------------
# the pieces of the step (duck-typed on the model: work on the twins above and on reference towers after install)
# ----------------------------------------------------------------------------------------------------------------
def embed_time(model, timestep: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``t = time_embedding(sinusoidal(timestep))``, ``t_mod = time_projection(t)`` (pipeline_mova.py:544-555):
    fp32 math, results rounded to the model dtype.  Returns ``(t [1, dim], t_mod [1, 6, dim])`` in bf16."""
    if timestep.numel() != 1:
        raise NotImplementedError("one timestep per call (MOVA runs B = 1)")
    memo = _memo(model, "_mova_b200_time", 2)
    te, tp = model.time_embedding, model.time_projection
    key = _tensor_key(timestep) + param_sig(te[0].weight, te[0].bias, te[2].weight, te[2].bias, tp[1].weight, tp[1].bias)
    hit = memo.get(key)
    if hit is not None:
        return hit
    dev = te[0].weight.device
    ts = timestep.detach().reshape(1).to(device=dev, dtype=torch.float32)
    s = ops.sinusoidal_embedding(model.freq_dim, ts)
    h = ops.gemv_f32(s, te[0].weight, te[0].bias, post_silu=True)
    t, t_bf = ops.gemv_f32(h, te[2].weight, te[2].bias, want_bf16=True)
    _, tm_bf = ops.gemv_f32(t, tp[1].weight, tp[1].bias, pre_silu=True, want_bf16=True)
    return memo.put(key, (timestep,), (t_bf.view(1, model.dim), tm_bf.view(1, 6, model.dim)))


def embed_text(model, context: torch.Tensor) -> torch.Tensor:
    """``text_embedding(context)`` (pipeline_mova.py:558-559): Linear + GELU-tanh + Linear on the 512 prompt tokens.
    Memoised per context tensor; the result is flagged so each block may memoise its text k / v on it."""
    memo = _memo(model, "_mova_b200_text", 4)
    te = model.text_embedding
    key = _tensor_key(context) + param_sig(te[0].weight, te[0].bias, te[2].weight, te[2].bias)
    hit = memo.get(key)
    if hit is not None:
        return hit
    ctx = context.to(device=te[0].weight.device, dtype=torch.bfloat16)
    h = ops.linear(ctx, te[0].weight, te[0].bias, epilogue=ops.EPI_GELU_TANH)
    out = ops.linear(h, te[2].weight, te[2].bias)
    setattr(out, _STATIC_FLAG, True)
    return memo.put(key, (context,), out)


def patchify(model, latents: torch.Tensor):
    """``model.patchify`` (wan_video_dit.py:399-409; wan_audio_dit.py:180-189): ``[1, C, F, H, W]`` (``[1, C, F]``)
    latents, fp32 or bf16 -> tokens ``[1, L, dim]`` bf16 and the token grid."""
    conv = model.patch_embedding
    w = conv.weight
    w2 = w.view(w.shape[0], -1) if w.is_contiguous() else w.reshape(w.shape[0], -1)
    psize = tuple(conv.kernel_size)
    cols = []
    for b in range(latents.shape[0]):  # B = 2 for a merged CFG pair (pipeline_mova.py:443-445), else 1
        x = latents[b]
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.to(torch.float32)
        x = x.to(w.device).contiguous()
        cols.append(ops.patchify(x, psize))
    cols = cols[0].unsqueeze(0) if len(cols) == 1 else torch.stack(cols)
    tokens = ops.linear(cols, w2, conv.bias)
    if x.dim() == 2:
        grid = (x.shape[1] // psize[0],)
    else:
        grid = (x.shape[1] // psize[0], x.shape[2] // psize[1], x.shape[3] // psize[2])
    return tokens, grid


def token_freqs(model, grid, device) -> torch.Tensor:
    """Per-token complex RoPE table as pipeline_mova.py:563-585 assembles it, memoised per (grid, device)."""
    memo = _memo(model, "_mova_b200_freqs", 4)
    key = (tuple(int(g) for g in grid), str(device))
    hit = memo.get(key)
    if hit is not None:
        return hit
    if len(grid) == 1:
        out = rope.audio_freqs(model.freqs, int(grid[0]), device)
    else:
        out = rope.video_freqs(model.freqs, tuple(int(g) for g in grid), device)
    return memo.put(key, (), out.contiguous())


def _head_tokens(model, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    head = model.head
    if isinstance(head, Head):
        return head(x, t)
    return Head.from_reference(head)(x, t)  # reference tower whose head was not swapped: same parameters


def head_unpatchify(model, x: torch.Tensor, t: torch.Tensor, grid, rows_per_rank: Optional[Sequence[int]] = None,
                    group=None) -> torch.Tensor:
    """``unpatchify(head(x, t))`` (pipeline_mova.py:603-607).  With ``rows_per_rank`` the tokens in ``x`` are this
    rank's chunk: the head runs on them and the all-gather moves the ``out_dim * prod(patch)`` wide head output
    instead of the hidden states."""
    y = _head_tokens(model, x, t)
    if rows_per_rank is not None:
        y = cpmod.all_gather_cat(y, rows_per_rank, group, dim=1)
    psize = tuple(model.patch_embedding.kernel_size)
    out_ch = y.shape[-1] // math.prod(psize)
    return torch.stack([ops.unpatchify(y[b], grid, psize, out_ch) for b in range(y.shape[0])])


# ----------------------------------------------------------------------------------------------------------------
# the step
# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def inference_single_step(self, visual_dit, visual_latents: torch.Tensor, audio_latents: Optional[torch.Tensor],
                          context: torch.Tensor, timestep: torch.Tensor, audio_timestep: Optional[torch.Tensor],
                          video_fps: float, cp_mesh=None):
    """Drop-in for ``MOVA.inference_single_step`` (pipeline_mova.py:500-609), same arguments and return value:
    ``(visual_output [1, 16, F, H/8, W/8], audio_output [1, 128, L_a])`` in bf16.  ``self`` needs ``audio_dit``,
    ``dual_tower_bridge`` and ``forward_dual_tower_dit`` (a ``MOVA`` pipeline after ``dualforce_b200.install``).

    ``context [B, 512, text_dim]`` with B = 2 runs the CFG pair as ONE batched forward (what the reference's
    ``cfg_merge`` branch, pipeline_mova.py:443-445, expects back: outputs ``[2, ...]``, positive first): the two
    samples share the timestep, so they are extra rows of every GEMM / LayerNorm and the batch dimension of every
    attention launch; latents may be given once (``[1, ...]``) or per sample.  Single GPU only (cp_mesh=None)."""
    from . import pipeline as pl

    audio_dit = self.audio_dit
    pre = getattr(self, "_pre_forward", None)
    if pre is not None:  # accelerate offload hooks (pipeline_mova.py:496-498, 525-527)
        pre(visual_dit)
        pre(audio_dit)
        pre(self.dual_tower_bridge)
    if audio_timestep is None:
        audio_timestep = timestep

    visual_t, visual_t_mod = embed_time(visual_dit, timestep)
    audio_t, audio_t_mod = embed_time(audio_dit, audio_timestep)
    visual_context = embed_text(visual_dit, context)
    audio_context = embed_text(audio_dit, context)

    visual_x, grid_size = patchify(visual_dit, visual_latents)
    audio_x, (f,) = patchify(audio_dit, audio_latents)
    B = context.shape[0] if context.dim() == 3 else 1
    if B > 1:
        # merged CFG pair (the `cfg_merge` branch of pipeline_mova.py:443-445): ONE forward over [positive, negative]
        # prompts.  Latents given once are shared by the samples: patchified once, tokens repeated.
        if visual_x.shape[0] == 1:
            visual_x = visual_x.repeat(B, 1, 1)
        if audio_x.shape[0] == 1:
            audio_x = audio_x.repeat(B, 1, 1)
        if visual_x.shape[0] != B or audio_x.shape[0] != B:
            raise ValueError(f"inference_single_step: {B} prompts but {visual_x.shape[0]} / {audio_x.shape[0]} latents")
    visual_freqs = token_freqs(visual_dit, grid_size, visual_x.device)
    audio_freqs = token_freqs(audio_dit, (f,), audio_x.device)

    args = dict(visual_dit=visual_dit, visual_x=visual_x, audio_x=audio_x, visual_context=visual_context,
                audio_context=audio_context, visual_t_mod=visual_t_mod, audio_t_mod=audio_t_mod,
                visual_freqs=visual_freqs, audio_freqs=audio_freqs, grid_size=grid_size, video_fps=video_fps)
    rows = group = None
    if cp_mesh is not None and not getattr(self, "mova_b200_cuda_graph", False):
        # keep the video tokens sharded through the head: the only full-length tensor that crosses NVLink is [L_v, 64]
        visual_x, audio_x, rows, group = pl._forward_eager(self, **args, cp_mesh=cp_mesh, _gather=False)
    else:
        visual_x, audio_x = self.forward_dual_tower_dit(**args, cp_mesh=cp_mesh)

    visual_output = head_unpatchify(visual_dit, visual_x, visual_t, grid_size, rows, group)
    audio_output = head_unpatchify(audio_dit, audio_x, audio_t, (f,))
    return visual_output, audio_output


def guided_update(noise_pred_posi: torch.Tensor, noise_pred_nega: Optional[torch.Tensor], sample: torch.Tensor,
                  cfg_scale: float, sigma_from, sigma_to, *, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """What ``MOVA.__call__`` does with the two predictions of one scheduler iteration, in one kernel:
    ``noise = nega + cfg_scale * (posi - nega)`` (pipeline_mova.py:456-460; ``noise_pred_nega=None`` is the
    ``cfg_scale == 1`` branch, :439-441) followed by ``FlowMatchPairScheduler.step_from_to``
    (schedulers/flow_match_pair.py:213-227): ``sample + noise * (sigma_to - sigma_from)``.  The sigmas come from the
    reference's scheduler (``scheduler.timestep_to_sigma``; ``sigma_to = 0`` after the last step).  ``sample`` is the
    fp32 latent tensor; the result is fp32 and may be written in place (``out=sample``)."""
    dsigma = float(sigma_to) - float(sigma_from)
    posi = noise_pred_posi.contiguous()
    nega = noise_pred_nega.contiguous() if noise_pred_nega is not None else None
    if sample.dtype != torch.float32:
        raise TypeError(f"guided_update: latents are kept in fp32 by the pipeline (pipeline_mova.py:378-399), got {sample.dtype}")
    return ops.cfg_euler_step(posi, nega, sample, float(cfg_scale), dsigma, out=out)


@torch.no_grad()
def denoising_loop(pipe, latents: torch.Tensor, condition: torch.Tensor, audio_latents: torch.Tensor,
                   prompt_embeds: torch.Tensor, negative_prompt_embeds: Optional[torch.Tensor], paired_timesteps,
                   timestep_to_sigma, video_fps: float, cfg_scale: float = 5.0, cp_mesh=None, pick_visual_dit=None,
                   final_sigma: float = 0.0, cfg_merge: bool = False):
    """The diffusion loop of ``MOVA.__call__`` (pipeline_mova.py:405-487) on the B200 step: per scheduler iteration
    two ``inference_single_step`` calls (one when ``cfg_scale == 1``), then CFG + ``step_from_to`` fused in
    ``guided_update``.  ``latents [1, 16, F, H, W]`` / ``audio_latents [1, 128, L_a]`` are the fp32 noise tensors,
    ``condition [1, 20, F, H, W]`` the mask + first-frame channels (:416); ``paired_timesteps`` is the ``[N, 2]`` tensor
    of ``scheduler.get_pairs()`` and ``timestep_to_sigma`` the reference scheduler's own lookup (:198-211) -- the
    scheduler stays the reference's.  ``pick_visual_dit(timestep_value) -> model`` implements the high / low-noise
    expert switch (:407-413); default: ``pipe.video_dit``.  Returns the final fp32 ``(latents, audio_latents)``.

    The model input ``cat([latents, condition], dim=1)`` (:416) is one persistent buffer whose first 16 channels are
    updated in place by the fused kernel, so no concatenation copy runs per step.

    ``cfg_merge=True`` (the reference's ``cfg_merge`` switch, :348, :443-445): the positive and the negative prompt
    run as ONE batched ``inference_single_step`` (B = 2) per iteration instead of two calls; single GPU only."""
    dev = latents.device
    n_lat = latents.shape[1]
    model_input = torch.cat([latents.float(), condition.float()], dim=1).contiguous()  # once per video
    lat_view = model_input[:, :n_lat]
    audio = audio_latents.float().contiguous().clone()
    total = paired_timesteps.shape[0]
    both = None
    for idx in range(total):
        t_v, t_a = paired_timesteps[idx]
        visual_dit = pick_visual_dit(float(t_v)) if pick_visual_dit is not None else pipe.video_dit
        ts_v = t_v.reshape(1).to(device=dev, dtype=torch.float32)
        ts_a = t_a.reshape(1).to(device=dev, dtype=torch.float32)
        kw = dict(visual_dit=visual_dit, visual_latents=model_input, audio_latents=audio, timestep=ts_v,
                  audio_timestep=ts_a, video_fps=video_fps, cp_mesh=cp_mesh)
        neg_v = neg_a = None
        if cfg_merge and cfg_scale != 1.0:
            if both is None:
                both = torch.cat([prompt_embeds, negative_prompt_embeds], dim=0)  # once per video: keeps the memos warm
            out_v, out_a = pipe.inference_single_step(context=both, **kw)
            pos_v, neg_v = out_v[0:1], out_v[1:2]
            pos_a, neg_a = out_a[0:1], out_a[1:2]
        else:
            pos_v, pos_a = pipe.inference_single_step(context=prompt_embeds, **kw)
            if cfg_scale != 1.0:
                neg_v, neg_a = pipe.inference_single_step(context=negative_prompt_embeds, **kw)
        nxt = paired_timesteps[idx + 1] if idx + 1 < total else None
        sig_v, sig_a = timestep_to_sigma(t_v), timestep_to_sigma(t_a)
        sig_v_to = timestep_to_sigma(nxt[0]) if nxt is not None else final_sigma
        sig_a_to = timestep_to_sigma(nxt[1]) if nxt is not None else final_sigma
        guided_update(pos_v, neg_v, lat_view, cfg_scale, sig_v, sig_v_to, out=lat_view)
        guided_update(pos_a, neg_a, audio, cfg_scale, sig_a, sig_a_to, out=audio)
    return lat_view.clone(), audio


def bind(pipe) -> None:
    """Bind ``pipe.inference_single_step`` to the B200 step (called by ``dualforce_b200.install``)."""
    pipe.inference_single_step = types.MethodType(inference_single_step, pipe)
