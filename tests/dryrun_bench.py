#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- dry run of bench.py's B200 arm on a machine without a GPU.

    python tests/dryrun_bench.py [bench args, default: --video-layers 1 --audio-layers 1 --frames 5 --steps 1 --warmup 3]

Kernels are replaced by tests/emulated_ops.py, CUDA events by host timers, the device by the CPU; what runs is
bench.py's own control flow -- model construction, warm-up, timed regions, roofline bookkeeping, the forward-level and
step-level e2e legs, the CPU baseline and the assembly of the JSON line -- so that a typo in the harness cannot cost
the round's only hardware measurement.  The numbers it prints mean nothing."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import torch  # noqa: E402

import dualforce_b200  # noqa: E402
import emulated_ops  # noqa: E402


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class FakeStream:
    cuda_stream = 0

    def wait_event(self, ev):
        pass


def main():
    import bench

    real_device = torch.device

    class DeviceShim:
        """torch.device("cuda", i) -> cpu, everything else untouched (also usable as a context manager)."""

        def __new__(cls, *args, **kwargs):
            if args and (args[0] == "cuda" or (isinstance(args[0], real_device) and args[0].type == "cuda")):
                return real_device("cpu")
            return real_device(*args, **kwargs)

    bench_torch = torch
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.Event = FakeEvent
    torch.cuda.current_stream = lambda *a, **k: FakeStream()
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.device = DeviceShim
    dualforce_b200._lib.require_device = lambda index: None
    for name, fn in emulated_ops.ENTRY_POINTS.items():
        setattr(dualforce_b200.ops, name, fn)
    # the attention front end also feeds bench.py's per-launch timers: keep that bookkeeping alive in the emulation
    emu_attention = emulated_ops.attention

    def attention(q, k, v, num_heads, **kw):
        rec = dualforce_b200._lib._TIMERS
        e0 = FakeEvent()
        e0.record()
        out = emu_attention(q, k, v, num_heads, **kw)
        dualforce_b200._lib.LAUNCHES += 1
        if rec is not None:
            e1 = FakeEvent()
            e1.record()
            rec.append((e0, e1, q.shape[0], q.shape[1], k.shape[1], num_heads, q.shape[2] // num_heads))
        return out

    dualforce_b200.ops.attention = attention
    # multi-rank dry run (torchrun --nproc-per-node 2 tests/dryrun_bench.py --gpus 2 ...): gloo instead of NCCL
    import torch.distributed as dist
    import torch.distributed.device_mesh as device_mesh

    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **kw: real_init("gloo", **{k: v for k, v in kw.items() if k != "device_id"})
    real_mesh = device_mesh.init_device_mesh
    device_mesh.init_device_mesh = lambda device_type, *a, **kw: real_mesh("cpu", *a, **kw)
    del bench_torch
    sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["--video-layers", "1", "--audio-layers", "1", "--frames", "5",
                                                 "--steps", "1", "--warmup", "3"])
    bench.main()


if __name__ == "__main__":
    main()
