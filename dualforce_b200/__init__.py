"""dualforce_b200 -- B200-native (sm_100a) implementation of the MOVA dual-tower DiT denoising-block forward.

Public surface (mirrors the reference's module / attention-processor interface for this path):

* :func:`install` -- swap the B200 modules into an existing ``MOVA`` pipeline (``pipeline_mova.py:124-148`` idiom),
* :func:`forward_dual_tower_dit` -- drop-in for ``MOVA.forward_dual_tower_dit`` (``pipeline_mova.py:612-711``),
* :func:`inference_single_step` -- drop-in for ``MOVA.inference_single_step`` (``pipeline_mova.py:500-609``): the
  embeddings, patchify, head and unpatchify either side of the path, with the per-prompt work memoised,
* :class:`DiTBlock`, :class:`AttentionModule`, :class:`DualTowerConditionalBridge`, ... -- module twins,
* :mod:`dualforce_b200.ops` -- the kernels behind them (C ABI in ``include/mova_b200.h``).

The compute lives in ``libmova_b200.so`` (hand-written CUDA for sm_100a); importing this package does not need a
GPU, calling anything does -- there is no CPU or PyTorch fallback.
"""
from . import _lib, cp, ops, rope  # noqa: F401
from ._lib import MovaB200Error  # noqa: F401
from .modules import (AttentionModule, ConditionalCrossAttention, ConditionalCrossAttentionBlock,  # noqa: F401
                      CrossAttention, CrossModalInteractionController, DiTBlock, DualTowerConditionalBridge,
                      GateModule, RotaryEmbedding, SelfAttention, USPAttention)
from .pipeline import CPRuntime, forward_dual_tower_dit, install  # noqa: F401
from .step import Head, WanAudioModel, WanModel, inference_single_step  # noqa: F401

__version__ = "0.1.0"
