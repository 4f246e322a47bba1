"""-m gpu: the context-parallel path.  world_size 1 (one process, NCCL group of one) exercises the segmented GEMM /
RMSNorm layouts, the side-stream all-to-all plumbing and the LSE merge on any single-GPU box; world_size 2 runs when
two GPUs are visible."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import cp_check

        cp_check.run_check(rank, world)
    finally:
        dist.destroy_process_group()


def test_cp_world1_layouts():
    mp.spawn(_worker, args=(1, _free_port()), nprocs=1, join=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_cp_world2():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
