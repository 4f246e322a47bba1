"""B200 twins of the reference's hot-path modules: same constructor arguments, ``forward`` signatures, attribute
names and state-dict keys as

* ``mova/diffusion/models/wan_video_dit.py``: AttentionModule (:154-161), SelfAttention (:164-189),
  CrossAttention (:211-247), GateModule (:250-255), DiTBlock (:257-291);
* ``mova/diffusion/models/interactionv2.py``: ConditionalCrossAttention (:210-251),
  ConditionalCrossAttentionBlock (:315-350), DualTowerConditionalBridge (:357-593),

so they load the reference's checkpoints and drop into ``MOVA`` (``dualforce_b200.install``).  ``nn.Linear`` /
``nn.LayerNorm`` / ``nn.RMSNorm`` objects are kept only as parameter containers; every ``forward`` here runs the
sm_100a kernels of ``libmova_b200.so`` through :mod:`dualforce_b200.ops` -- there is no PyTorch compute path.

Fusion plan per DiTBlock (17 launches instead of ~60 library kernels):
  add_to_f32(modulation + t_mod) -> LN+modulate -> QKV GEMM (one launch, packed weight) -> RMSNorm+RoPE (q, k in
  place) -> attention on strided q/k/v views -> o-proj GEMM with ``x + gate * (.)`` epilogue -> LN(affine) ->
  q GEMM -> RMSNorm -> [text k/v GEMM + RMSNorm] -> attention -> o-proj GEMM with residual epilogue ->
  LN+modulate -> FFN-1 GEMM with GELU-tanh epilogue -> FFN-2 GEMM with gated-residual epilogue.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
from torch.nn import RMSNorm

from . import ops, rope

__all__ = [
    "AttentionModule", "USPAttention", "SelfAttention", "CrossAttention", "GateModule", "DiTBlock", "ConditionalCrossAttention",
    "ConditionalCrossAttentionBlock", "CrossModalInteractionController", "RotaryEmbedding",
    "DualTowerConditionalBridge", "merged_linear",
]


def _pack_linears(linears: List[nn.Linear]) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Concatenate the weights (and biases) of several nn.Linear along the output dim into one buffer and re-point
    each layer's parameters at views of it: one GEMM launch, no duplicated weight memory, state-dict keys intact."""
    w = torch.cat([l.weight.data for l in linears], dim=0).contiguous()
    b = None
    if all(l.bias is not None for l in linears):
        b = torch.cat([l.bias.data for l in linears], dim=0).contiguous()
    row = 0
    for l in linears:
        n = l.weight.shape[0]
        l.weight.data = w[row:row + n]
        if b is not None:
            l.bias.data = b[row:row + n]
        row += n
    return w, b


def merged_linear(layer: nn.Module) -> nn.Linear:
    """A plain ``nn.Linear`` for ``layer``.  The reference's LoRA wrapper (``LoRALinear``: ``original_layer``,
    ``lora_A``, ``lora_B``, ``scaling``; engine/trainer/accelerate/lora_utils.py:19-109) computes
    ``W x + scaling * B (A x)``; the kernels take one weight, so the adapter is folded in exactly as the reference's own
    ``merge_weights`` / ``MOVALoRA.merge_lora_weights`` does (lora_utils.py:95-109, pipelines/mova_lora.py:190-220):
    ``W' = W + scaling * B A`` -- evaluated in fp32 and rounded once to the weight dtype.  A plain ``nn.Linear`` is
    returned unchanged (same object, so its Parameters stay shared with the reference module)."""
    if isinstance(layer, nn.Linear):
        return layer
    if all(hasattr(layer, a) for a in ("original_layer", "lora_A", "lora_B", "scaling")):
        base = merged_linear(layer.original_layer)
        w = base.weight.data
        delta = (layer.lora_B.weight.data.float() @ layer.lora_A.weight.data.float()) * float(layer.scaling)
        out = nn.Linear(base.in_features, base.out_features, bias=base.bias is not None, device="meta")
        out.weight = nn.Parameter((w.float() + delta.to(w.device)).to(w.dtype), requires_grad=False)
        if base.bias is not None:
            out.bias = nn.Parameter(base.bias.data.clone(), requires_grad=False)
        return out
    raise TypeError(f"expected nn.Linear or a LoRA-wrapped linear, got {type(layer).__module__}.{type(layer).__name__}")


def _kv_splits(n_queries: int, n_keys: int, num_heads: int = 12, sms: int = 148) -> int:
    """Split-KV factor for occupancy-bound cross-attention: few queries against many keys (v2a: 403 audio queries x
    43 120 video keys x 12 heads = 48 CTAs of 128 query rows on 148 SMs, each walking 337 key blocks).  The keys are cut
    in ``s`` equal chunks run as the batch dimension of ONE launch (so ``s`` must divide the key count; 43 120 =
    2^4 x 5 x 7^2 x 11) and merged exactly with their log-sum-exps.  Picks the ``s <= 8`` that minimises
    waves x key blocks per CTA; 1 when the launch already fills the GPU or the chunks would get shorter than 2048 keys.
    Measured on B200 (round 1, 256-row tiles): 5 chunks beat the single launch."""
    if n_queries > 1024 or n_keys < 8192:
        return 1
    ctas = -(-n_queries // 128) * num_heads
    best, best_cost = 1, -(-ctas // sms) * -(-n_keys // 128)
    for s in range(2, 9):
        if n_keys % s or n_keys // s < 2048:
            continue
        cost = -(-ctas * s // sms) * -(-(n_keys // s) // 128) + 2  # + the merge pass
        if cost < best_cost:
            best, best_cost = s, cost
    return best


def param_sig(*tensors) -> tuple:
    """(pointer, version) of every tensor: a cache key that changes when a parameter is moved OR updated in place
    (``load_state_dict`` / ``copy_`` bump ``_version``; views of a packed buffer share the buffer's counter)."""
    sig = []
    for t in tensors:
        if t is None:
            sig.append(None)
            continue
        try:
            ver = t._version
        except RuntimeError:  # inference-mode tensor: immutable
            ver = -1
        sig.append((t.data_ptr(), ver))
    return tuple(sig)


class _Packed:
    """Lazily built packed weights, rebuilt if the owning parameters moved (``.to()``, ``.cuda()``, dtype cast)."""

    def __init__(self):
        self.w: Optional[torch.Tensor] = None
        self.b: Optional[torch.Tensor] = None
        self._sig = None

    def get(self, linears: List[nn.Linear]):
        sig = tuple((l.weight.data_ptr(), l.weight.dtype, str(l.weight.device)) for l in linears)
        if self.w is None or sig != self._sig:
            self.w, self.b = _pack_linears(linears)
            self._sig = tuple((l.weight.data_ptr(), l.weight.dtype, str(l.weight.device)) for l in linears)
        return self.w, self.b


def prepack(module: nn.Module) -> None:
    """Build every packed weight buffer below ``module`` NOW, on the current stream.  Packing frees the original
    parameter storage (the Linears are re-pointed at views of the packed buffer), so it must not happen lazily on a
    side stream while the allocating stream may recycle that storage: the context-parallel path calls this before it
    forks the audio stream (found on hardware: a cold first forward with the audio side stream corrupted weights)."""
    for m in module.modules():
        if isinstance(m, SelfAttention):
            m._qkv.get([m.q, m.k, m.v])
        elif isinstance(m, (CrossAttention, ConditionalCrossAttention)):
            m._kv.get([m.k, m.v])


class AttentionModule(nn.Module):
    """wan_video_dit.py:154-161 -- ``forward(q, k, v)`` on flat ``[B, S, H*D]`` tensors."""

    def __init__(self, num_heads: int):
        super().__init__()
        self.num_heads = num_heads

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
        splits = _kv_splits(q.shape[1], k.shape[1], self.num_heads) if q.shape[0] == 1 else 1
        if splits > 1:
            return split_kv_attention(q, k, v, self.num_heads, splits)
        return ops.attention(q, k, v, self.num_heads)


def split_kv_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int, splits: int) -> torch.Tensor:
    """Few queries, many keys (v2a: 403 audio queries x 43 120 video keys x 12 heads = 48 CTAs on 148 SMs): cut the
    keys in ``splits`` equal chunks, run them as the BATCH dimension of one attention launch (48 x splits CTAs) and
    merge the partial results exactly with their log-sum-exps."""
    _, Sq, HD = q.shape
    chunk = k.shape[1] // splits
    kc = k[0].unflatten(0, (splits, chunk))  # strided view [splits, chunk, HD] of the fused k|v buffer
    vc = v[0].unflatten(0, (splits, chunk))
    qe = q.expand(splits, Sq, HD).contiguous()  # the kernel's tensor map wants a real batch stride (6 MB at 360p)
    o, lse = ops.attention(qe, kc, vc, num_heads, return_lse=True)
    return ops.lse_merge(o, lse, num_heads).unsqueeze(0)


class USPAttention(nn.Module):
    """wan_video_dit.py:192-208 -- the context-parallel attention processor ``MOVA.replace_attention`` installs:
    ``forward(q, k, v)`` on sequence shards ``[B, S/P, H*D]``, Ulysses head <-> sequence all-to-all over ``group``
    (default: the world group) around the sm_100a attention kernel.  ``attn_type`` is accepted for signature
    compatibility and ignored (there is one kernel).  ``dualforce_b200.install`` does not need this class -- its
    ``forward_dual_tower_dit`` shards only the video tower and exchanges without pack copies -- it exists for callers
    that keep the reference loop and swap processors one by one."""

    def __init__(self, num_heads: int, attn_type=None, group=None):
        super().__init__()
        self.num_heads = num_heads
        self.attn_type = attn_type
        self.group = group

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
        from . import cp

        return cp.ulysses_attention(q, k, v, self.num_heads, ops.attention, self.group)


class SelfAttention(nn.Module):
    """wan_video_dit.py:164-189."""

    def __init__(self, dim: int, num_heads: int, eps: float = 1e-6):
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(dim, dim)
        self.v = nn.Linear(dim, dim)
        self.o = nn.Linear(dim, dim)
        self.norm_q = RMSNorm(dim, eps=eps)
        self.norm_k = RMSNorm(dim, eps=eps)
        self.attn = AttentionModule(self.num_heads)
        self._qkv = _Packed()

    def qkv(self, x: torch.Tensor, freqs) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """q, k, v after projection, full-width RMSNorm and interleaved RoPE, as column views of one buffer."""
        d = self.dim
        w, b = self._qkv.get([self.q, self.k, self.v])
        qkv = ops.linear(x, w, b)
        cos, sin = rope.as_tables(freqs)
        q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
        ops.rmsnorm_rope_(q, self.norm_q.weight, self.norm_q.eps, head_dim=self.head_dim, cos=cos, sin=sin,
                          rope_mode=ops.ROPE_INTERLEAVED)
        ops.rmsnorm_rope_(k, self.norm_k.weight, self.norm_k.eps, head_dim=self.head_dim, cos=cos, sin=sin,
                          rope_mode=ops.ROPE_INTERLEAVED)
        return q, k, v

    def forward(self, x: torch.Tensor, freqs) -> torch.Tensor:
        q, k, v = self.qkv(x, freqs)
        return ops.linear(self.attn(q, k, v), self.o.weight, self.o.bias)


class CrossAttention(nn.Module):
    """wan_video_dit.py:211-247 (text cross-attention; MOVA ships ``has_image_input=False``)."""

    def __init__(self, dim: int, num_heads: int, eps: float = 1e-6, has_image_input: bool = False):
        super().__init__()
        if has_image_input:
            raise NotImplementedError(
                "CrossAttention(has_image_input=True) (CLIP image branch, wan_video_dit.py:233-246) is disabled in the "
                "MOVA checkpoints and not built here")
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(dim, dim)
        self.v = nn.Linear(dim, dim)
        self.o = nn.Linear(dim, dim)
        self.norm_q = RMSNorm(dim, eps=eps)
        self.norm_k = RMSNorm(dim, eps=eps)
        self.has_image_input = has_image_input
        self.attn = AttentionModule(self.num_heads)
        self._kv = _Packed()

    def kv(self, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """k = RMSNorm(Linear(y)), v = Linear(y) of the text context.  The prompt embedding is constant over the
        denoising schedule, so for a context flagged static by ``step.embed_text`` the result is memoised (the
        reference recomputes it in every layer of every forward, wan_video_dit.py:218-223)."""
        d = self.dim
        w, b = self._kv.get([self.k, self.v])
        static = getattr(y, "_mova_b200_static", False)
        if static:
            try:
                version = y._version
            except RuntimeError:  # inference-mode tensor
                version = -1
            key = (y.data_ptr(), tuple(y.shape), version) + param_sig(w, b, self.norm_k.weight)
            memo = self.__dict__.setdefault("_kv_memo", {})
            hit = memo.get(key)
            if hit is not None:
                return hit[1]
        kv = ops.linear(y, w, b)
        k, v = kv[..., :d], kv[..., d:]
        ops.rmsnorm_rope_(k, self.norm_k.weight, self.norm_k.eps)
        if static:
            while len(memo) >= 4:  # positive + negative prompt, with slack
                memo.pop(next(iter(memo)))
            memo[key] = (y, (k, v))  # y kept alive so its storage is not recycled under the key
        return k, v

    def attend(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Everything before the output projection."""
        q = ops.linear(x, self.q.weight, self.q.bias)
        ops.rmsnorm_rope_(q, self.norm_q.weight, self.norm_q.eps)
        k, v = self.kv(y)
        return self.attn(q, k, v)

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return ops.linear(self.attend(x, y), self.o.weight, self.o.bias)


class GateModule(nn.Module):
    """wan_video_dit.py:250-255.  Kept for structural parity; the gate is fused into the GEMM epilogues."""

    def forward(self, x, gate, residual):
        raise RuntimeError("GateModule is fused into the o-proj / FFN-2 GEMM epilogues and is never called")


class DiTBlock(nn.Module):
    """wan_video_dit.py:257-291 -- ``forward(x, context, t_mod, freqs)``."""

    def __init__(self, has_image_input: bool, dim: int, num_heads: int, ffn_dim: int, eps: float = 1e-6):
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.ffn_dim = ffn_dim
        self.self_attn = SelfAttention(dim, num_heads, eps)
        self.cross_attn = CrossAttention(dim, num_heads, eps, has_image_input=has_image_input)
        self.norm1 = nn.LayerNorm(dim, eps=eps, elementwise_affine=False)
        self.norm2 = nn.LayerNorm(dim, eps=eps, elementwise_affine=False)
        self.norm3 = nn.LayerNorm(dim, eps=eps)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate="tanh"), nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)
        self.gate = GateModule()

    @classmethod
    def from_reference(cls, ref: nn.Module) -> "DiTBlock":
        """Wrap a reference ``DiTBlock`` without copying: the new block owns the very same Parameter objects."""
        if getattr(ref.cross_attn, "has_image_input", False):
            raise NotImplementedError("has_image_input=True blocks are not supported (off in the MOVA checkpoints)")
        blk = cls.__new__(cls)
        nn.Module.__init__(blk)
        blk.dim, blk.num_heads, blk.ffn_dim = ref.dim, ref.num_heads, ref.ffn_dim
        eps = ref.norm1.eps
        blk.self_attn = SelfAttention.__new__(SelfAttention)
        nn.Module.__init__(blk.self_attn)
        sa, rsa = blk.self_attn, ref.self_attn
        sa.dim, sa.num_heads, sa.head_dim = rsa.dim, rsa.num_heads, rsa.head_dim
        sa.q, sa.k, sa.v, sa.o = (merged_linear(m) for m in (rsa.q, rsa.k, rsa.v, rsa.o))
        sa.norm_q, sa.norm_k = rsa.norm_q, rsa.norm_k
        sa.attn = AttentionModule(rsa.num_heads)
        sa._qkv = _Packed()
        blk.cross_attn = CrossAttention.__new__(CrossAttention)
        nn.Module.__init__(blk.cross_attn)
        ca, rca = blk.cross_attn, ref.cross_attn
        ca.dim, ca.num_heads, ca.head_dim = rca.dim, rca.num_heads, rca.head_dim
        ca.q, ca.k, ca.v, ca.o = (merged_linear(m) for m in (rca.q, rca.k, rca.v, rca.o))
        ca.norm_q, ca.norm_k = rca.norm_q, rca.norm_k
        ca.has_image_input = False
        ca.attn = AttentionModule(rca.num_heads)
        ca._kv = _Packed()
        blk.norm1, blk.norm2, blk.norm3 = ref.norm1, ref.norm2, ref.norm3
        blk.ffn = ref.ffn
        blk.modulation = ref.modulation
        blk.gate = GateModule()
        assert eps == ref.norm2.eps
        return blk

    def modulation_f32(self, t_mod: torch.Tensor) -> torch.Tensor:
        """``modulation + t_mod`` as fp32 ``[6, d]`` rows: shift/scale/gate (msa), shift/scale/gate (mlp)."""
        if t_mod.dim() == 4:
            raise NotImplementedError("per-token t_mod (Wan2.2-5B, wan_video_dit.py:276-285) is not used by MOVA")
        return ops.add_to_f32(self.modulation.to(dtype=torch.bfloat16), t_mod.to(torch.bfloat16)).reshape(6, self.dim)

    def forward(self, x: torch.Tensor, context: torch.Tensor, t_mod: torch.Tensor, freqs) -> torch.Tensor:
        if x.shape[0] != 1 and t_mod.shape[0] != 1:  # per-sample modulation: run the samples one by one
            return torch.cat([self.forward(x[i:i + 1], context[i:i + 1], t_mod[i:i + 1], freqs)
                              for i in range(x.shape[0])], dim=0)
        # B > 1 with ONE t_mod (the CFG pair of pipeline_mova.py:443-445: same latents and timestep, two prompts): the
        # samples are extra rows of every GEMM / LayerNorm and the batch dimension of the attention launches
        mod = self.modulation_f32(t_mod)
        eps = self.norm1.eps
        sa, ca = self.self_attn, self.cross_attn
        # x = x + gate_msa * self_attn(modulate(norm1(x), shift_msa, scale_msa), freqs)
        h = ops.layernorm(x, eps, shift=mod[0], scale=mod[1])
        q, k, v = sa.qkv(h, freqs)
        a = sa.attn(q, k, v)
        x = ops.linear(a, sa.o.weight, sa.o.bias, epilogue=ops.EPI_RESIDUAL, residual=x, gate=mod[2])
        # x = x + cross_attn(norm3(x), context)
        h = ops.layernorm(x, self.norm3.eps, weight=self.norm3.weight, bias=self.norm3.bias, out=h)
        a = ca.attend(h, context)
        ops.linear(a, ca.o.weight, ca.o.bias, epilogue=ops.EPI_RESIDUAL, residual=x, out=x)
        # x = x + gate_mlp * ffn(modulate(norm2(x), shift_mlp, scale_mlp))
        ops.layernorm(x, self.norm2.eps, shift=mod[3], scale=mod[4], out=h)
        u = ops.linear(h, self.ffn[0].weight, self.ffn[0].bias, epilogue=ops.EPI_GELU_TANH)
        ops.linear(u, self.ffn[2].weight, self.ffn[2].bias, epilogue=ops.EPI_RESIDUAL, residual=x, gate=mod[5], out=x)
        return x


# ----------------------------------------------------------------------------------------------------------------
# bridge
# ----------------------------------------------------------------------------------------------------------------
class RotaryEmbedding(nn.Module):
    """interactionv2.py:12-37 -- cos/sin tables ``[B, L, dim]`` for arbitrary (fractional) positions."""

    def __init__(self, base: float, dim: int, device=None):
        super().__init__()
        self.base = base
        self.dim = dim
        self.attention_scaling = 1.0
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).to(device=device, dtype=torch.float) / dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)
        self.original_inv_freq = self.inv_freq
        self.rope_precision = "fp32"

    ROPE_PRECISIONS = ("fp32", "reference_bf16")

    @torch.no_grad()
    def forward(self, x, position_ids):
        """Two precisions, selected by ``self.rope_precision`` (``install(pipe, bridge_rope=...)``):

        ``"fp32"`` -- exact frequencies and tables: what the reference computes when it runs in fp32 (the CPU oracle).
        ``"reference_bf16"`` -- what the reference computes when it runs the way it is shipped: ``from_pretrained(...,
        torch_dtype=bfloat16)`` casts the non-persistent ``inv_freq`` buffer to bf16 (interactionv2.py:21-23), which
        perturbs every frequency by up to 2^-9 (0.8 rad at audio position 400), and ``build_aligned_freqs`` is called
        with the model dtype, so cos / sin are rounded to bf16 as well (:36, pipeline_mova.py:641-648).  Training
        (mova_train.py:914) and inference both run that way, so a MOVA checkpoint has only ever seen the rounded
        tables -- this mode reproduces them bit for bit (pinned by tests/golden/bridge_rope_bf16.npz).

        The frequencies are recomputed from the formula on every (memoised) call instead of read from the buffer, so
        the result does not depend on what ``.to(dtype)`` did to the module."""
        if self.rope_precision not in self.ROPE_PRECISIONS:
            raise ValueError(f"rope_precision must be one of {self.ROPE_PRECISIONS}, got {self.rope_precision!r}")
        inv_freq = 1.0 / (self.base ** (torch.arange(0, self.dim, 2, dtype=torch.int64).to(
            device=x.device, dtype=torch.float) / self.dim))
        if self.rope_precision == "reference_bf16":
            inv_freq = inv_freq.to(torch.bfloat16).float()
        inv = inv_freq[None, :, None].float().expand(position_ids.shape[0], -1, 1)
        pos = position_ids[:, None, :].float()
        freqs = (inv.float() @ pos.float()).transpose(1, 2)
        emb = torch.cat((freqs, freqs), dim=-1)
        cos = emb.cos() * self.attention_scaling
        sin = emb.sin() * self.attention_scaling
        if self.rope_precision == "reference_bf16":
            cos, sin = cos.to(torch.bfloat16).float(), sin.to(torch.bfloat16).float()
        return cos.to(dtype=x.dtype), sin.to(dtype=x.dtype)


class CrossModalInteractionController:
    """interactionv2.py:128-207 -- which layers exchange information."""

    def __init__(self, visual_layers: int = 30, audio_layers: int = 30):
        self.visual_layers = visual_layers
        self.audio_layers = audio_layers
        self.min_layers = min(visual_layers, audio_layers)

    def get_interaction_layers(self, strategy: str = "shallow_focus") -> Dict[str, List[Tuple[int, int]]]:
        n = self.min_layers
        if strategy == "shallow_focus":
            layers = list(range(0, min(10, n // 3)))
        elif strategy == "distributed":
            layers = list(range(0, n, 3))
        elif strategy == "progressive":
            layers = list(range(0, min(8, n)))
            if n > 8:
                layers = layers + list(range(8, n, 3))
        elif strategy == "custom":
            layers = [i for i in [0, 2, 4, 6, 8, 12, 16, 20] if i < n]
        elif strategy == "full":
            layers = list(range(0, n))
        else:
            raise ValueError(f"Unknown interaction strategy: {strategy}")
        return {"v2a": [(i, i) for i in layers], "a2v": [(i, i) for i in layers]}

    def should_interact(self, layer_idx: int, direction: str, interaction_mapping: Dict) -> bool:
        if direction not in interaction_mapping:
            return False
        return any(src == layer_idx for src, _ in interaction_mapping[direction])


class ConditionalCrossAttention(nn.Module):
    """interactionv2.py:210-251 -- q from the primary tower (un-normalised x), k/v from the conditioning tower,
    rotate-half RoPE on q (x_freqs) and k (y_freqs)."""

    def __init__(self, dim: int, kv_dim: int, num_heads: int, eps: float = 1e-6):
        super().__init__()
        self.q_dim = dim
        self.kv_dim = kv_dim
        self.num_heads = num_heads
        self.head_dim = self.q_dim // num_heads
        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(kv_dim, dim)
        self.v = nn.Linear(kv_dim, dim)
        self.o = nn.Linear(dim, dim)
        self.norm_q = RMSNorm(dim, eps=eps)
        self.norm_k = RMSNorm(dim, eps=eps)
        self.attn = AttentionModule(self.num_heads)
        self._kv = _Packed()

    def _rope_args(self, freqs):
        if freqs is None:
            return dict(rope_mode=ops.ROPE_NONE)
        cos, sin = rope.as_tables(freqs)
        return dict(cos=cos, sin=sin, rope_mode=ops.ROPE_HALF)

    def project_q(self, x, x_freqs):
        q = ops.linear(x, self.q.weight, self.q.bias)
        ops.rmsnorm_rope_(q, self.norm_q.weight, self.norm_q.eps, head_dim=self.head_dim, **self._rope_args(x_freqs))
        return q

    def project_kv(self, y, y_freqs):
        d = self.q_dim
        w, b = self._kv.get([self.k, self.v])
        kv = ops.linear(y, w, b)
        k, v = kv[..., :d], kv[..., d:]
        ops.rmsnorm_rope_(k, self.norm_k.weight, self.norm_k.eps, head_dim=self.head_dim, **self._rope_args(y_freqs))
        return k, v

    def attend(self, x, y, x_freqs=None, y_freqs=None):
        q = self.project_q(x, x_freqs)
        k, v = self.project_kv(y, y_freqs)
        return self.attn(q, k, v)  # AttentionModule: splits the keys when the launch would not fill the GPU

    def forward(self, x: torch.Tensor, y: torch.Tensor, x_freqs=None, y_freqs=None) -> torch.Tensor:
        return ops.linear(self.attend(x, y, x_freqs, y_freqs), self.o.weight, self.o.bias)


class ConditionalCrossAttentionBlock(nn.Module):
    """interactionv2.py:315-350 -- LayerNorm(affine) on the conditioning input, then ConditionalCrossAttention."""

    def __init__(self, dim: int, kv_dim: int, num_heads: int, eps: float = 1e-6, pooled_adaln: bool = False):
        super().__init__()
        if pooled_adaln:
            raise NotImplementedError(
                "pooled_adaln (PerFrameAttentionPooling + AdaLayerNorm, interactionv2.py:75-125,255-312) is off in the "
                "MOVA checkpoints and not built here")
        self.y_norm = nn.LayerNorm(kv_dim, eps=eps)
        self.inner = ConditionalCrossAttention(dim=dim, kv_dim=kv_dim, num_heads=num_heads, eps=eps)
        self.pooled_adaln = pooled_adaln

    @classmethod
    def from_reference(cls, ref: nn.Module) -> "ConditionalCrossAttentionBlock":
        if getattr(ref, "pooled_adaln", False):
            raise NotImplementedError("pooled_adaln conditioners are not supported (off in the MOVA checkpoints)")
        blk = cls.__new__(cls)
        nn.Module.__init__(blk)
        blk.y_norm = ref.y_norm
        blk.pooled_adaln = False
        inner = ConditionalCrossAttention.__new__(ConditionalCrossAttention)
        nn.Module.__init__(inner)
        r = ref.inner
        inner.q_dim, inner.kv_dim, inner.num_heads, inner.head_dim = r.q_dim, r.kv_dim, r.num_heads, r.head_dim
        inner.q, inner.k, inner.v, inner.o = (merged_linear(m) for m in (r.q, r.k, r.v, r.o))
        inner.norm_q, inner.norm_k = r.norm_q, r.norm_k
        inner.attn = AttentionModule(r.num_heads)
        inner._kv = _Packed()
        blk.inner = inner
        return blk

    def normed_condition(self, y: torch.Tensor) -> torch.Tensor:
        return ops.layernorm(y, self.y_norm.eps, weight=self.y_norm.weight, bias=self.y_norm.bias)

    def forward(self, x, y, x_freqs=None, y_freqs=None, video_grid_size=None) -> torch.Tensor:
        return self.inner(x=x, y=self.normed_condition(y), x_freqs=x_freqs, y_freqs=y_freqs)

    def forward_residual(self, x, y, x_freqs, y_freqs, scale: float) -> torch.Tensor:
        """``x + scale * block(x, y)`` with the residual fused into the o-proj epilogue (interactionv2.py:535)."""
        a = self.inner.attend(x, self.normed_condition(y), x_freqs, y_freqs)
        return ops.linear(a, self.inner.o.weight, self.inner.o.bias, epilogue=ops.EPI_RESIDUAL, residual=x,
                          scale=float(scale))


class DualTowerConditionalBridge(nn.Module):
    """interactionv2.py:357-593 -- bidirectional video<->audio conditioning, both directions read the pre-bridge
    states.  Constructor arguments and defaults follow the reference; the MOVA-360p checkpoint uses
    visual 5120 / audio 1536 / head_dim 128 / "full" / apply_cross_rope / audio_fps 50."""

    _repeated_blocks = ("ConditionalCrossAttentionBlock",)

    def __init__(self, visual_layers: int = 30, audio_layers: int = 30, visual_hidden_dim: int = 3072,
                 audio_hidden_dim: int = 1536, audio_fps: float = 44100.0 / 2048.0, head_dim: int = 128,
                 interaction_strategy: str = "shallow_focus", apply_cross_rope: bool = False,
                 apply_first_frame_bias_in_rope: bool = False, trainable_condition_scale: bool = False,
                 pooled_adaln: bool = False):
        super().__init__()
        if head_dim != 128:
            raise NotImplementedError("the sm_100a attention kernel is built for head_dim 128 (MOVA's value)")
        self.visual_hidden_dim = visual_hidden_dim
        self.audio_hidden_dim = audio_hidden_dim
        self.audio_fps = audio_fps
        self.head_dim = head_dim
        self.apply_cross_rope = apply_cross_rope
        self.apply_first_frame_bias_in_rope = apply_first_frame_bias_in_rope
        self.trainable_condition_scale = trainable_condition_scale
        self.pooled_adaln = pooled_adaln
        if trainable_condition_scale:
            self.condition_scale = nn.Parameter(torch.tensor([1.0], dtype=torch.float32))
        else:
            self.condition_scale = 1.0
        self.controller = CrossModalInteractionController(visual_layers, audio_layers)
        self.interaction_mapping = self.controller.get_interaction_layers(interaction_strategy)
        self.audio_to_video_conditioners = nn.ModuleDict()
        self.video_to_audio_conditioners = nn.ModuleDict()
        self.rotary = RotaryEmbedding(base=10000.0, dim=head_dim)
        for v_layer, _ in self.interaction_mapping["a2v"]:
            self.audio_to_video_conditioners[str(v_layer)] = ConditionalCrossAttentionBlock(
                dim=visual_hidden_dim, kv_dim=audio_hidden_dim, num_heads=visual_hidden_dim // head_dim,
                pooled_adaln=False)
        for a_layer, _ in self.interaction_mapping["v2a"]:
            self.video_to_audio_conditioners[str(a_layer)] = ConditionalCrossAttentionBlock(
                dim=audio_hidden_dim, kv_dim=visual_hidden_dim, num_heads=audio_hidden_dim // head_dim,
                pooled_adaln=self.pooled_adaln)
        self._freq_cache: Dict[tuple, tuple] = {}

    @classmethod
    def from_reference(cls, ref: nn.Module) -> "DualTowerConditionalBridge":
        br = cls.__new__(cls)
        nn.Module.__init__(br)
        for name in ("visual_hidden_dim", "audio_hidden_dim", "audio_fps", "head_dim", "apply_cross_rope",
                     "apply_first_frame_bias_in_rope", "trainable_condition_scale", "pooled_adaln", "controller",
                     "interaction_mapping"):
            setattr(br, name, getattr(ref, name))
        if br.head_dim != 128:
            raise NotImplementedError("the sm_100a attention kernel is built for head_dim 128 (MOVA's value)")
        br.condition_scale = ref.condition_scale
        br.rotary = RotaryEmbedding(base=10000.0, dim=br.head_dim)  # fp32-stable twin, see RotaryEmbedding.forward
        br.audio_to_video_conditioners = nn.ModuleDict(
            {k: ConditionalCrossAttentionBlock.from_reference(m) for k, m in ref.audio_to_video_conditioners.items()})
        br.video_to_audio_conditioners = nn.ModuleDict(
            {k: ConditionalCrossAttentionBlock.from_reference(m) for k, m in ref.video_to_audio_conditioners.items()})
        br._freq_cache = {}
        return br

    @torch.no_grad()
    def build_aligned_freqs(self, video_fps: float, grid_size: Tuple[int, int, int], audio_steps: int,
                            device: Optional[torch.device] = None, dtype: Optional[torch.dtype] = None):
        """interactionv2.py:420-475.  Audio token i sits at position i; every token of video frame f sits at
        ``f * audio_fps / (video_fps / 4)`` (VAE temporal stride 4).  Returns ``((cos_v, sin_v), (cos_a, sin_a))``
        of shape ``[1, L, head_dim]``.  The result only depends on the arguments, so it is memoised (the reference
        rebuilds it in every forward, pipeline_mova.py:641-648)."""
        f_v, h, w = (int(g) for g in grid_size)
        L_a = int(audio_steps)
        device = torch.device(device) if device is not None else next(self.parameters()).device
        dtype = dtype or torch.float32
        key = (float(video_fps), f_v, h, w, L_a, str(device), dtype, float(self.audio_fps),
               bool(self.apply_first_frame_bias_in_rope), self.rotary.rope_precision)
        hit = self._freq_cache.get(key)
        if hit is not None:
            return hit
        audio_pos = torch.arange(L_a, device=device, dtype=torch.float32).unsqueeze(0)
        if self.apply_first_frame_bias_in_rope:
            eff = float(video_fps) / 4.0
            t_starts = torch.zeros((f_v,), device=device, dtype=torch.float32)
            if f_v > 1:
                t_starts[1:] = (1.0 / float(video_fps)) + torch.arange(f_v - 1, device=device,
                                                                       dtype=torch.float32) * (1.0 / eff)
            per_frame = t_starts * float(self.audio_fps)
        else:
            scale = float(self.audio_fps) / float(video_fps / 4.0)
            per_frame = torch.arange(f_v, device=device, dtype=torch.float32) * scale
        video_pos = per_frame.repeat_interleave(h * w).unsqueeze(0)
        dummy_v = torch.zeros((1, 1, 1), device=device, dtype=dtype)
        cos_v, sin_v = self.rotary(dummy_v, position_ids=video_pos)
        cos_a, sin_a = self.rotary(dummy_v, position_ids=audio_pos)
        out = ((cos_v, sin_v), (cos_a, sin_a))
        if len(self._freq_cache) >= 4:
            self._freq_cache.pop(next(iter(self._freq_cache)))
        self._freq_cache[key] = out
        return out

    @property
    def bridge_rope(self) -> str:
        """Precision of the aligned cross-RoPE tables: ``"fp32"`` or ``"reference_bf16"`` (RotaryEmbedding.forward)."""
        return self.rotary.rope_precision

    @bridge_rope.setter
    def bridge_rope(self, value: str) -> None:
        if value not in RotaryEmbedding.ROPE_PRECISIONS:
            raise ValueError(f"bridge_rope must be one of {RotaryEmbedding.ROPE_PRECISIONS}, got {value!r}")
        self.rotary.rope_precision = value

    def should_interact(self, layer_idx: int, direction: str) -> bool:
        return self.controller.should_interact(layer_idx, direction, self.interaction_mapping)

    def _scale(self, condition_scale) -> float:
        scale = condition_scale if condition_scale is not None else self.condition_scale
        if isinstance(scale, torch.Tensor):
            scale = float(scale.detach().float().reshape(-1)[0].item())
        return float(scale)

    def apply_conditional_control(self, layer_idx: int, direction: str, primary_hidden_states: torch.Tensor,
                                  condition_hidden_states: torch.Tensor, x_freqs=None, y_freqs=None,
                                  condition_scale: Optional[float] = None, video_grid_size=None) -> torch.Tensor:
        """interactionv2.py:480-537: ``primary + conditioner(primary, condition) * scale``."""
        if not self.should_interact(layer_idx, direction):
            return primary_hidden_states
        if direction == "a2v":
            conditioner = self.audio_to_video_conditioners[str(layer_idx)]
        elif direction == "v2a":
            conditioner = self.video_to_audio_conditioners[str(layer_idx)]
        else:
            raise ValueError(f"Invalid direction: {direction}")
        return conditioner.forward_residual(primary_hidden_states, condition_hidden_states, x_freqs, y_freqs,
                                            self._scale(condition_scale))

    def forward(self, layer_idx: int, visual_hidden_states: torch.Tensor, audio_hidden_states: torch.Tensor, *,
                x_freqs=None, y_freqs=None, a2v_condition_scale: Optional[float] = None,
                v2a_condition_scale: Optional[float] = None, condition_scale: Optional[float] = None,
                video_grid_size=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """interactionv2.py:539-593."""
        visual_conditioned = self.apply_conditional_control(
            layer_idx, "a2v", visual_hidden_states, audio_hidden_states, x_freqs=x_freqs, y_freqs=y_freqs,
            condition_scale=a2v_condition_scale if a2v_condition_scale is not None else condition_scale,
            video_grid_size=video_grid_size)
        audio_conditioned = self.apply_conditional_control(
            layer_idx, "v2a", audio_hidden_states, visual_hidden_states, x_freqs=y_freqs, y_freqs=x_freqs,
            condition_scale=v2a_condition_scale if v2a_condition_scale is not None else condition_scale,
            video_grid_size=video_grid_size)
        return visual_conditioned, audio_conditioned
