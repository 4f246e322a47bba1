#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- dry run of the `-m gpu` test CODE on a machine without a GPU.

    python tests/dryrun_gpu_tests.py [pytest args, default: the step tests]

The kernels are replaced by tests/emulated_ops.py and `.cuda()` / `device="cuda"` are mapped to the CPU, so what runs
is the tests' own logic -- shapes, oracle calls, tolerances -- against the product's real host code.  It exists to
catch mistakes in GPU tests that were written when no GPU time was left (their first real run is then a measurement of
the kernels, not of the test code).  It proves nothing about the CUDA kernels and is not collected by pytest."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import pytest  # noqa: E402
import torch  # noqa: E402

import dualforce_b200  # noqa: E402
import emulated_ops  # noqa: E402


def _cpu_device(d):
    if isinstance(d, str) and d.startswith("cuda"):
        return "cpu"
    if isinstance(d, torch.device) and d.type == "cuda":
        return torch.device("cpu")
    return d


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.is_available = lambda: True  # conftest.py then leaves the gpu-marked tests in
    torch.cuda.synchronize = lambda *a, **k: None
    dualforce_b200._lib.require_device = lambda index: None
    for name, fn in emulated_ops.ENTRY_POINTS.items():
        setattr(dualforce_b200.ops, name, fn)
    orig_to = torch.nn.Module.to

    def module_to(self, *args, **kwargs):
        args = tuple(_cpu_device(a) for a in args)
        if "device" in kwargs:
            kwargs["device"] = _cpu_device(kwargs["device"])
        return orig_to(self, *args, **kwargs)

    torch.nn.Module.to = module_to
    orig_tensor_to = torch.Tensor.to

    def tensor_to(self, *args, **kwargs):
        args = tuple(_cpu_device(a) for a in args)
        if "device" in kwargs:
            kwargs["device"] = _cpu_device(kwargs["device"])
        return orig_tensor_to(self, *args, **kwargs)

    torch.Tensor.to = tensor_to

    def cpu_factory(fn):
        def wrapped(*args, **kwargs):
            if "device" in kwargs:
                kwargs["device"] = _cpu_device(kwargs["device"])
            return fn(*args, **kwargs)
        return wrapped

    for name in ("tensor", "ones", "zeros", "empty", "randn", "arange", "full"):
        setattr(torch, name, cpu_factory(getattr(torch, name)))
    args = sys.argv[1:] or [os.path.join(HERE, "test_gpu_step.py"), "-k", "not context_parallel"]
    return pytest.main(["-q", "--runxfail", "-m", "gpu", "-p", "no:cacheprovider", *args])


if __name__ == "__main__":
    sys.exit(main())
