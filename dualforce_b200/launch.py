"""Activation hook: run an unmodified reference script on the B200 path.

    torchrun --nproc-per-node 8 -m dualforce_b200.launch /path/to/reference/scripts/inference_single.py --cp_size 8 ...

``activate()`` wraps ``MOVA.__call__`` (mova/diffusion/pipelines/pipeline_mova.py:322) so that the first call of a
pipeline object runs ``dualforce_b200.install(pipe)`` -- by then ``from_pretrained`` and ``pipe.to(device)`` /
``enable_*_offload`` (scripts/inference_single.py:77-97) have happened -- and then proceeds with the reference's own
``__call__``, which reaches ``inference_single_step`` / ``forward_dual_tower_dit`` through the instance attributes that
``install`` bound.  Nothing in the reference tree is edited; ``MOVA_B200_CUDA_GRAPH=1`` selects graph replay.
"""
from __future__ import annotations

import functools
import os
import runpy
import sys


def activate(pipeline_cls=None):
    """Patch ``pipeline_cls.__call__`` (default: the reference's ``MOVA``) to install the B200 modules on first use.
    Returns the patched class."""
    if pipeline_cls is None:
        from mova.diffusion.pipelines.pipeline_mova import MOVA as pipeline_cls  # the reference package must be importable
    if getattr(pipeline_cls, "_mova_b200_activated", False):
        return pipeline_cls
    original = pipeline_cls.__call__

    @functools.wraps(original)
    def call(self, *args, **kwargs):
        if not self.__dict__.get("_mova_b200_installed", False):
            from . import install

            install(self, cuda_graph=os.environ.get("MOVA_B200_CUDA_GRAPH", "0") == "1")
            self.__dict__["_mova_b200_installed"] = True
        return original(self, *args, **kwargs)

    pipeline_cls.__call__ = call
    pipeline_cls._mova_b200_activated = True
    return pipeline_cls


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m dualforce_b200.launch <reference script.py> [script args...]")
    activate()
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
