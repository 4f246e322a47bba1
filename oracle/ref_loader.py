"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference modules from /root/reference on CPU.

Used in the authoring container by ``oracle/make_golden.py`` (to generate the committed fixtures under
``tests/golden/``) and by ``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent, e.g. on
the GPU box).  Nothing in the product package imports this file.

The reference needs four harness-side shims to import here; none touches its arithmetic:
  1. ``diffusers`` is not installed      -> stub ConfigMixin / register_to_config / ModelMixin,
  2. ``yunchang`` is not installed       -> stub with an AttnType enum that has ``FA`` (wan_video_dit.py:193),
  3. ``flash_attn`` has no CPU backend   -> hidden, so flash_attention() takes its SDPA branch (:85-90),
  4. ``@torch.compile`` cannot build on CPU here -> TORCHDYNAMO_DISABLE=1 (eager, same formulas).
"""
from __future__ import annotations

import ast
import enum
import os
import sys
import types
from typing import Optional

REFERENCE_ROOT = os.environ.get("MOVA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mova", "diffusion", "models"))


_loaded = None


def load():
    """Returns a namespace with the reference's wan_video_dit, wan_audio_dit, interactionv2 and functional modules
    plus ``forward_dual_tower_dit`` (pipeline_mova.py:612-711) and ``inference_single_step`` (:500-609) lifted, source
    unchanged, out of pipeline_mova.py."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    import torch
    import torch.nn as nn

    # (1) diffusers stubs
    if "diffusers" not in sys.modules:
        diffusers = types.ModuleType("diffusers")
        cfg = types.ModuleType("diffusers.configuration_utils")
        mu = types.ModuleType("diffusers.models.modeling_utils")
        models = types.ModuleType("diffusers.models")

        class ConfigMixin:  # noqa: D401
            pass

        def register_to_config(fn):
            return fn

        class ModelMixin(nn.Module):
            @property
            def dtype(self):
                return next(self.parameters()).dtype

            @property
            def device(self):
                return next(self.parameters()).device

        cfg.ConfigMixin, cfg.register_to_config = ConfigMixin, register_to_config
        mu.ModelMixin = ModelMixin
        diffusers.configuration_utils, diffusers.models = cfg, models
        models.modeling_utils = mu
        sys.modules.update({"diffusers": diffusers, "diffusers.configuration_utils": cfg, "diffusers.models": models,
                            "diffusers.models.modeling_utils": mu})
    # (2) yunchang stub
    if "yunchang" not in sys.modules:
        yc = types.ModuleType("yunchang")
        yk = types.ModuleType("yunchang.kernels")

        class AttnType(enum.Enum):
            FA = "fa"
            FA3 = "fa3"
            TORCH = "torch"

        yk.AttnType = AttnType
        yc.LongContextAttention = None
        yc.kernels = yk
        sys.modules.update({"yunchang": yc, "yunchang.kernels": yk})
    # (3) hide flash_attn / FA3 / sage so the SDPA branch is taken
    for name in ("flash_attn", "flash_attn_interface", "kernels", "sageattention"):
        sys.modules[name] = None
    # import the three model files by package path without running mova/__init__.py side effects
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for pkg in ("mova", "mova.diffusion", "mova.diffusion.models", "mova.distributed"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REFERENCE_ROOT, *pkg.split("."))]
            sys.modules[pkg] = m
    import importlib

    wan_video_dit = importlib.import_module("mova.diffusion.models.wan_video_dit")
    wan_audio_dit = importlib.import_module("mova.diffusion.models.wan_audio_dit")
    interactionv2 = importlib.import_module("mova.diffusion.models.interactionv2")
    functional = importlib.import_module("mova.distributed.functional")

    # lift MOVA.forward_dual_tower_dit and MOVA.inference_single_step verbatim (pipeline_mova.py itself needs
    # diffusers/transformers/ftfy to import)
    src_path = os.path.join(REFERENCE_ROOT, "mova", "diffusion", "pipelines", "pipeline_mova.py")
    with open(src_path) as f:
        tree = ast.parse(f.read())
    ns = {
        "torch": torch, "Optional": Optional, "DeviceMesh": object,
        "sinusoidal_embedding_1d": wan_video_dit.sinusoidal_embedding_1d,
        "_sp_split_tensor": functional._sp_split_tensor, "_sp_split_tensor_dim_0": functional._sp_split_tensor_dim_0,
        "_sp_all_gather_avg": functional._sp_all_gather_avg,
    }
    lifted_lines = {}
    for name in ("forward_dual_tower_dit", "inference_single_step"):
        fn_node = None
        for node in ast.walk(tree):
            if isinstance(node, ast.FunctionDef) and node.name == name:
                fn_node = node
                break
        assert fn_node is not None, f"{name} not found in pipeline_mova.py"
        fn_node.decorator_list = []
        module = ast.Module(body=[fn_node], type_ignores=[])
        ast.fix_missing_locations(module)
        exec(compile(module, src_path, "exec"), ns)
        lifted_lines[name] = (fn_node.lineno, fn_node.end_lineno)

    _loaded = types.SimpleNamespace(
        wan_video_dit=wan_video_dit, wan_audio_dit=wan_audio_dit, interactionv2=interactionv2, functional=functional,
        forward_dual_tower_dit=ns["forward_dual_tower_dit"], inference_single_step=ns["inference_single_step"],
        lines=lifted_lines["forward_dual_tower_dit"], step_lines=lifted_lines["inference_single_step"])
    return _loaded
