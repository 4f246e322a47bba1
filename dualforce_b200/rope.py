"""RoPE table plumbing for the two conventions of the reference.

* tower self-attention (wan_video_dit.py:106-137, wan_audio_dit.py:48-60): the pipeline hands every block one
  complex128 table ``freqs [L, 1, 64]`` (pipeline_mova.py:563-585); the kernels want fp32 ``cos/sin [L, 64]``.
* bridge cross-attention (interactionv2.py:12-72, 420-475): ``(cos, sin)`` pairs ``[1, L, 128]`` in the model
  dtype; the kernels want fp32 ``[L, 128]``.

The reference rebuilds and re-casts these tables in every layer of every forward (K12 in SURVEY.md); here each
distinct table is converted once and memoised on the identity of the source tensor.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

# key -> (source tensors, (cos, sin)).  The sources are kept alive on purpose: while an entry exists its
# storage cannot be recycled for a different table, so (data_ptr, shape, version) identifies the content.
_CACHE: Dict[tuple, tuple] = {}
_CACHE_MAX = 16


_ENABLED = True


def set_cache(enabled: bool) -> None:
    """Disable the memo while a CUDA graph is being captured: the conversions must become graph nodes that re-run on
    every replay (the static input buffers keep their address but not their content)."""
    global _ENABLED
    _ENABLED = bool(enabled)


def _key(t: torch.Tensor) -> tuple:
    try:
        version = t._version
    except RuntimeError:  # inference-mode tensors carry no version counter (and cannot be modified in place)
        version = -1
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, version, str(t.device))


def _remember(key: tuple, sources: tuple, value):
    if not _ENABLED:
        return value
    if len(_CACHE) >= _CACHE_MAX:
        _CACHE.pop(next(iter(_CACHE)))
    _CACHE[key] = (sources, value)
    return value


def clear_cache() -> None:
    _CACHE.clear()


def tables_from_complex(freqs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """complex ``[L, 1, 64]`` (or ``[L, 64]``) -> contiguous fp32 ``(cos [L, 64], sin [L, 64])``."""
    if not torch.is_complex(freqs):
        raise TypeError(f"expected a complex RoPE table, got {freqs.dtype}")
    key = ("c",) + _key(freqs)
    hit = _CACHE.get(key) if _ENABLED else None
    if hit is not None:
        return hit[1]
    f = freqs.reshape(freqs.shape[0], -1)
    cos = f.real.to(torch.float32).contiguous()
    sin = f.imag.to(torch.float32).contiguous()
    return _remember(key, (freqs,), (cos, sin))


def tables_from_cos_sin(cos: torch.Tensor, sin: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``[1, L, 128]`` (any float dtype) -> contiguous fp32 ``[L, 128]`` pair."""
    key = ("r",) + _key(cos) + _key(sin)
    hit = _CACHE.get(key) if _ENABLED else None
    if hit is not None:
        return hit[1]
    src = (cos, sin)
    if cos.dim() == 3:
        if cos.shape[0] != 1:
            raise NotImplementedError("per-sample bridge RoPE tables (batch > 1) are not used by MOVA")
        cos, sin = cos[0], sin[0]
    return _remember(key, src, (cos.to(torch.float32).contiguous(), sin.to(torch.float32).contiguous()))


def as_tables(freqs) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """Accept what the reference passes (complex tensor, or a (cos, sin) pair) or already converted tables."""
    if freqs is None:
        return None
    if isinstance(freqs, (tuple, list)):
        cos, sin = freqs
        if cos.dtype == torch.float32 and cos.dim() == 2 and cos.is_contiguous() and sin.is_contiguous():
            return cos, sin
        return tables_from_cos_sin(cos, sin)
    try:  # DTensor shards (wan_video_dit.py:183-184)
        from torch.distributed.tensor import DTensor

        if isinstance(freqs, DTensor):
            freqs = freqs.to_local()
    except Exception:  # pragma: no cover
        pass
    return tables_from_complex(freqs)


def precompute_freqs_cis(dim: int, end: int = 1024, theta: float = 10000.0, s: float = 1.0) -> torch.Tensor:
    """1-D complex table, same formula as wan_video_dit.py:115-121 (float64 angles); ``s`` scales the positions
    (wan_audio_dit.py:53-60)."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].double() / dim))
    freqs = torch.outer(torch.arange(end, dtype=torch.float64) * s, freqs)
    return torch.polar(torch.ones_like(freqs), freqs)


def precompute_freqs_cis_3d(dim: int, end: int = 1024, theta: float = 10000.0):
    """(frame, height, width) tables with the 22/21/21 complex-pair split of wan_video_dit.py:106-112."""
    return (precompute_freqs_cis(dim - 2 * (dim // 3), end, theta), precompute_freqs_cis(dim // 3, end, theta),
            precompute_freqs_cis(dim // 3, end, theta))


def precompute_freqs_cis_1d(dim: int, end: int = 16384, theta: float = 10000.0):
    """Audio tower: one 1-D table over all 64 pairs, chunked in 3 (wan_audio_dit.py:48-50)."""
    return precompute_freqs_cis(dim, end, theta).chunk(3, dim=-1)


def legacy_precompute_freqs_cis_1d(dim: int, end: int = 16384, theta: float = 10000.0, base_tps: float = 4.0,
                                   target_tps: float = 44100 / 2048):
    """``vae_type="oobleck"`` tables (wan_audio_dit.py:37-45): rescaled positions on the first 22 pairs, identity
    rotation on the other 2 x 21 (MOVA ships ``vae_type="dac"``; kept for constructor parity)."""
    s = float(base_tps) / float(target_tps)
    f = precompute_freqs_cis(dim - 2 * (dim // 3), end, theta, s)
    ones = torch.ones_like(precompute_freqs_cis(dim // 3, end, theta, s))
    return f, ones, ones


def video_freqs(tables, grid_size, device) -> torch.Tensor:
    """Per-token complex table ``[f*h*w, 1, 64]`` exactly as pipeline_mova.py:563-570 assembles it."""
    f, h, w = grid_size
    t = tuple(x.to(device) for x in tables)
    return torch.cat([
        t[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
        t[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
        t[2][:w].view(1, 1, w, -1).expand(f, h, w, -1),
    ], dim=-1).reshape(f * h * w, 1, -1)


def audio_freqs(tables, length: int, device) -> torch.Tensor:
    """Per-token complex table ``[L_a, 1, 64]`` as in pipeline_mova.py:577-585."""
    t = tuple(x.to(device) for x in tables)
    return torch.cat([t[0][:length], t[1][:length], t[2][:length]], dim=-1).reshape(length, 1, -1)
