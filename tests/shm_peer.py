"""CPU stand-in for dualforce_b200.peer.CudaIpcWindow (test infrastructure): the same window interface over files in
/dev/shm that every rank of a gloo test maps, so dualforce_b200.peer.PeerExchange -- offsets, flag indices, epochs,
ragged rows -- and the pipeline's peer branch run on CPU with world_size > 1.  Copies are synchronous; the flag wait is a
polling loop with a time-out."""
import atexit
import os
import time

import numpy as np
import torch
import torch.distributed as dist

FLAG_WORDS = 1024


class ShmWindow:
    def __init__(self, rank: int, size: int, tag: str, group=None, jitter_ms: float = 0.0):
        self.rank, self.size, self.tag, self.group = rank, size, tag, group
        # de-synchronise the ranks: a random pause before every push and wait.  The exchange sends no acknowledgements
        # (dualforce_b200/peer.py relies on data dependencies), so a rank that runs ahead must never overwrite what a
        # slower peer has not consumed yet
        self.jitter_ms = jitter_ms
        self._rng = np.random.default_rng(1234 + rank)
        self.capacity = 0
        self.maps = []
        self.generation = 0
        self.pushes = 0
        self.waits = 0
        atexit.register(self._unlink)

    def _path(self, r: int, gen: int) -> str:
        return f"/dev/shm/mova_peer_{self.tag}_{gen}_{r}"

    def _unlink(self):
        for gen in range(self.generation + 1):
            try:
                os.unlink(self._path(self.rank, gen))
            except OSError:
                pass

    def ensure(self, nbytes: int) -> None:
        if nbytes <= self.capacity:
            return
        dist.barrier(group=self.group)
        self.maps = []
        self.generation += 1
        want = (int(nbytes) + 4095) & ~4095
        own = np.memmap(self._path(self.rank, self.generation), dtype=np.uint8, mode="w+", shape=(want,))
        own[:] = 0
        own.flush()
        dist.barrier(group=self.group)
        for r in range(self.size):
            self.maps.append(own if r == self.rank else
                             np.memmap(self._path(r, self.generation), dtype=np.uint8, mode="r+", shape=(want,)))
        self.capacity = want
        dist.barrier(group=self.group)

    def _bytes(self, r: int) -> torch.Tensor:
        return torch.from_numpy(self.maps[r])

    def local_tensor(self, offset, shape, dtype):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        assert offset % 16 == 0 and offset + n <= self.capacity
        return self._bytes(self.rank)[offset:offset + n].view(dtype).view(*shape)

    def _pause(self):
        if self.jitter_ms > 0:
            time.sleep(float(self._rng.uniform(0, self.jitter_ms)) * 1e-3)

    def push(self, copies, flags, epoch, stream=None):
        self._pause()
        for r, off, t in copies:
            assert t.is_contiguous()
            src = t.reshape(-1).view(torch.uint8)
            assert off + src.numel() <= self.capacity
            self._bytes(r)[off:off + src.numel()].copy_(src)
        for r, idx in flags:
            assert 0 <= idx < FLAG_WORDS
            self.maps[r][8 * idx:8 * idx + 8].view(np.uint64)[0] = np.uint64(epoch)
        self.pushes += 1

    def wait(self, first_flag, n_flags, epoch, stream=None):
        self._pause()
        words = self.maps[self.rank][8 * first_flag:8 * (first_flag + n_flags)].view(np.uint64)
        t0 = time.time()
        while not bool((words >= np.uint64(epoch)).all()):
            if time.time() - t0 > 60:
                raise TimeoutError(f"rank {self.rank}: flags {first_flag}+{n_flags} < epoch {epoch}: {words.tolist()}")
            time.sleep(0.0005)
        self.waits += 1
