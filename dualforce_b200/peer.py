"""Ulysses head <-> sequence exchange over NVSwitch peer memory (``csrc/peer.cu``).

Replaces, on the data path, the two ``dist.all_to_all_single`` calls per video self-attention that stand in for
yunchang's LongContextAttention (``USPAttention.forward``, mova/diffusion/models/wan_video_dit.py:192-208).

Every rank owns one *window* of device memory that its peers map (CUDA IPC) and write into:

    [ flag words : 8 KB ][ recv : G x L x 3w bf16 ][ back : G x cp x rows(rank) x w bf16 ]

* in  (q|k|v of head group g, my tokens, heads of rank d)  ->  rank d's ``recv[g, rows_before(me) : +rows(me), :]``
* out (attention output of head group g, tokens of rank d)  ->  rank d's ``back[g, me, :, :]``

Both are contiguous chunks (the QKV GEMM writes destination-rank-major, the attention output is token-major), so one
exchange is ``cp`` plain copies on the copy engines plus one flag word per destination; the consumer's stream polls its
own flag words before the attention / o-projection launch.  Nothing on this path needs an SM, which is the point: an
NCCL all-to-all launched beside the attention kernel waits for SMs the attention kernel holds (csrc/peer.cu).

Ordering invariants the layout relies on (no acknowledgements are sent):
  * a rank pushes the q|k|v of layer l+1 only after its own o-projection of layer l, which needed the attention output
    of layer l from EVERY peer -- so every peer has finished reading its ``recv`` of layer l;
  * ``back`` of layer l is read by the o-projection, which precedes the QKV GEMM of layer l+1 in stream order;
  * between two forwards there is at least one collective (the final all-gather of the hidden states / latents),
    so a change of shapes -- hence of offsets -- never races with a reader; growing the window is itself collective
    (device synchronise + barrier on every rank before the old mapping is dropped).
Flag words carry a monotonically increasing epoch (one per exchange round, the same sequence on every rank).
"""
from __future__ import annotations

import ctypes
import warnings
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["PeerExchange", "CudaIpcWindow", "PeerUnavailable"]

FLAG_BYTES = 8192
MAX_GROUPS = 32
WAIT_TIMEOUT_MS = 120_000
EPOCH_SRC_SLOTS = 8
EPOCH_SRC_OFFSET = FLAG_BYTES - 8 * EPOCH_SRC_SLOTS  # words of the own window the epoch is staged in (memop signalling)


class PeerUnavailable(RuntimeError):
    """The peer-memory window could not be set up on every rank (no CUDA IPC / no peer access)."""


def _stream_ptr(stream) -> int:
    if stream is None:
        stream = torch.cuda.current_stream()
    return int(stream.cuda_stream)


class _RawCuda:
    """Minimal ``__cuda_array_interface__`` holder: lets torch view memory this package allocated itself."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def deal_by_destination(copies, flags, k: int):
    """Split one push into at most ``k`` pushes, destinations dealt round-robin in order of first appearance; every
    destination's copies and flag stay together and in order.  Returns ``[(copies_j, flags_j), ...]`` (non-empty)."""
    dests = []
    for r in [c[0] for c in copies] + [f[0] for f in flags]:
        if r not in dests:
            dests.append(r)
    out = []
    for j in range(min(k, len(dests))):
        mine = set(dests[j::k])
        out.append(([c for c in copies if c[0] in mine], [f for f in flags if f[0] in mine]))
    return out


class CudaIpcWindow:
    """One cudaMalloc'd window per rank, mapped into every peer of ``group`` through CUDA IPC.

    ``signal``: how flag words are written and awaited -- ``"memops"`` (stream memory operations + 8-byte copy-engine
    copies: no kernel, so nothing queues behind an attention kernel that owns every SM) or ``"kernel"`` (a 32-thread
    store kernel and a polling kernel with a time-out trap); ``None`` picks memops when the driver offers them."""

    signal = None

    def __init__(self, group, rank: int, size: int, device: torch.device):
        from . import _lib

        self._lib = _lib
        with torch.cuda.device(device):
            supported = bool(_lib.load().mova_b200_peer_memops_supported())
        if self.signal not in (None, "memops", "kernel"):
            raise ValueError(f"CudaIpcWindow.signal must be None, 'memops' or 'kernel', got {self.signal!r}")
        if self.signal == "memops" and not supported:
            raise PeerUnavailable("stream memory operations requested but not supported by this driver / device")
        self.memops = supported if self.signal is None else (self.signal == "memops")
        self.group, self.rank, self.size, self.device = group, rank, size, device
        self.capacity = 0
        self.local_ptr = 0
        self.ptrs: List[int] = []  # base address of every rank's window as seen from this device
        self._base: Optional[torch.Tensor] = None

    # -- collective ------------------------------------------------------------------------------------------------
    def ensure(self, nbytes: int) -> None:
        """Make every rank's window at least ``nbytes`` large.  Collective whenever it has to (re)allocate -- all ranks
        call it with the same sizes in the same order, so they all take the same branch."""
        if nbytes <= self.capacity:
            return
        lib = self._lib.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody is still writing into, or reading from, the old windows
        self._release()
        want = (int(nbytes) + (1 << 21) - 1) & ~((1 << 21) - 1)
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            rc = lib.mova_b200_peer_alloc(want, ctypes.byref(ptr), handle)
        err = "" if rc == 0 else self._lib.last_error()
        gathered = [None] * self.size
        dist.all_gather_object(gathered, (rc, bytes(handle.raw), err), group=self.group)
        bad = [(r, g[2]) for r, g in enumerate(gathered) if g[0] != 0]
        if bad:
            if rc == 0:
                lib.mova_b200_peer_free(ptr)
            raise PeerUnavailable(f"peer window allocation failed on ranks {bad}")
        ptrs, opened, fail = [0] * self.size, [], ""
        with torch.cuda.device(self.device):
            for r in range(self.size):
                if r == self.rank:
                    ptrs[r] = int(ptr.value)
                    continue
                p = ctypes.c_void_p()
                if lib.mova_b200_peer_open(gathered[r][1], ctypes.byref(p)) != 0:
                    fail = f"rank {self.rank} cannot map rank {r}: {self._lib.last_error()}"
                    break
                ptrs[r] = int(p.value)
                opened.append(int(p.value))
        status = [None] * self.size
        dist.all_gather_object(status, fail, group=self.group)
        if any(status):
            with torch.cuda.device(self.device):
                for p in opened:
                    lib.mova_b200_peer_close(ctypes.c_void_p(p))
                lib.mova_b200_peer_free(ptr)
            raise PeerUnavailable("; ".join(s for s in status if s))
        self.local_ptr, self.ptrs, self.capacity = int(ptr.value), ptrs, want
        self._base = torch.as_tensor(_RawCuda(self.local_ptr, want), device=self.device)
        if self._base.data_ptr() != self.local_ptr:
            raise PeerUnavailable("torch copied the window instead of aliasing it")
        dist.barrier(group=self.group)

    def _release(self) -> None:
        if not self.capacity:
            return
        lib = self._lib.load()
        with torch.cuda.device(self.device):
            for r, p in enumerate(self.ptrs):
                if r != self.rank:
                    lib.mova_b200_peer_close(ctypes.c_void_p(p))
            self._base = None
            lib.mova_b200_peer_free(ctypes.c_void_p(self.local_ptr))
        self.capacity, self.local_ptr, self.ptrs = 0, 0, []

    def close(self) -> None:
        if self.capacity:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            self._release()

    # -- local -----------------------------------------------------------------------------------------------------
    def local_tensor(self, offset: int, shape: Sequence[int], dtype: torch.dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        assert offset % 16 == 0 and offset + nbytes <= self.capacity
        return self._base[offset:offset + nbytes].view(dtype).view(*shape)

    def push(self, copies: Sequence[Tuple[int, int, torch.Tensor]], flags: Sequence[Tuple[int, int]], epoch: int,
             stream=None) -> None:
        """``copies``: (destination rank, byte offset in its window, contiguous source tensor on this device);
        ``flags``: (destination rank, flag index) stored with ``epoch`` after the copies.

        ``stream`` may be a list of streams: the destinations are then dealt round-robin over them (a destination's
        copies and its flag stay on one stream, in order), so the ~9 us a copy-engine operation costs on a stream are
        paid in parallel.  The caller orders the streams against producers and consumers with events."""
        if isinstance(stream, (list, tuple)) and len(stream) > 1:
            for j, (cj, fj) in enumerate(deal_by_destination(copies, flags, len(stream))):
                self._push_one(cj, fj, epoch, stream[j], j)
            return
        if isinstance(stream, (list, tuple)):
            stream = stream[0] if stream else None
        self._push_one(copies, flags, epoch, stream, 0)

    def _push_one(self, copies, flags, epoch: int, stream, slot: int) -> None:
        n, m = len(copies), len(flags)
        dst = (ctypes.c_void_p * max(n, 1))()
        src = (ctypes.c_void_p * max(n, 1))()
        nb = (ctypes.c_int64 * max(n, 1))()
        for i, (r, off, t) in enumerate(copies):
            assert t.is_contiguous() and t.device == self.device
            nbytes = t.numel() * t.element_size()
            assert off + nbytes <= self.capacity
            dst[i], src[i], nb[i] = self.ptrs[r] + off, t.data_ptr(), nbytes
        fl = (ctypes.c_void_p * max(m, 1))()
        local_mask = 0
        for i, (r, idx) in enumerate(flags):
            assert 0 <= idx < EPOCH_SRC_OFFSET // 8
            fl[i] = self.ptrs[r] + 8 * idx
            if r == self.rank:
                local_mask |= 1 << i
        assert 0 <= slot < EPOCH_SRC_SLOTS  # one staging word per push stream: their epochs may differ in flight
        epoch_src = (self.local_ptr + EPOCH_SRC_OFFSET + 8 * slot) if self.memops else None
        rc = self._lib.load().mova_b200_peer_push(n, dst, src, nb, m, fl, local_mask, int(epoch), epoch_src,
                                                  _stream_ptr(stream))
        self._lib.check(rc, "mova_b200_peer_push", launches=0 if self.memops else 1)

    def wait(self, first_flag: int, n_flags: int, epoch: int, stream=None) -> None:
        assert 0 <= first_flag and first_flag + n_flags <= EPOCH_SRC_OFFSET // 8
        rc = self._lib.load().mova_b200_peer_wait(self.local_ptr + 8 * first_flag, n_flags, int(epoch), WAIT_TIMEOUT_MS,
                                                  1 if self.memops else 0, _stream_ptr(stream))
        self._lib.check(rc, "mova_b200_peer_wait", launches=0 if self.memops else 1)


class PeerExchange:
    """Offsets, flag indices and epochs of the exchange; the window (``CudaIpcWindow`` on the device, a shared-memory
    stand-in in the CPU tests) moves the bytes."""

    def __init__(self, window, rank: int, size: int):
        self.window, self.rank, self.size = window, rank, size
        self.epoch = 0
        self._shape = None

    @staticmethod
    def flag_index(direction: int, group: int, src: int, cp: int) -> int:
        return (direction * MAX_GROUPS + group) * cp + src

    def begin(self, G: int, L: int, C: int, rows_per_rank: Sequence[int], w: int,
              dtype: torch.dtype = torch.bfloat16) -> Tuple[torch.Tensor, torch.Tensor]:
        """Start one exchange round (one self-attention): returns this rank's ``recv [G, L, C]`` and
        ``back [G, cp, rows(me), w]`` views of its window.  ``C`` = 3 w for the fused q|k|v buffer."""
        cp = self.size
        if G > MAX_GROUPS or 2 * MAX_GROUPS * cp * 8 > EPOCH_SRC_OFFSET:
            raise ValueError(f"peer exchange: {G} head groups x {cp} ranks exceed the flag area")
        assert len(rows_per_rank) == cp and sum(rows_per_rank) == L
        item = torch.empty((), dtype=dtype).element_size()
        self._item = item
        self._rows = [int(r) for r in rows_per_rank]
        self._row_off = [sum(self._rows[:r]) for r in range(cp)]
        self._G, self._L, self._C, self._w = G, L, C, w
        self._recv_off = FLAG_BYTES
        self._back_off = (self._recv_off + G * L * C * item + 1023) & ~1023
        total = self._back_off + G * cp * max(self._rows) * w * item
        self.window.ensure(total)
        self.epoch += 1
        recv = self.window.local_tensor(self._recv_off, (G, L, C), dtype)
        back = self.window.local_tensor(self._back_off, (G, cp, self._rows[self.rank], w), dtype)
        return recv, back

    # Remote and local chunks go separately.  Remote chunks are peer copies: copy engines, they run beside a kernel
    # that owns every SM.  The chunk a rank keeps for itself is a same-device copy, which the driver may run on the
    # SMs -- queued behind the attention kernel it would stall the stream it sits on (measured,
    # profiles/r02_ce_overlap_probe.json), so local chunks are pushed where no attention kernel is in the way: all
    # groups at once BEFORE the first remote push (inbound), and on the compute stream after the last attention set
    # (outbound).
    def _order(self):
        cp, me = self.size, self.rank
        return [(me + 1 + i) % cp for i in range(cp - 1)]  # every rank starts with a different peer

    def push_in(self, g: int, send_g: torch.Tensor, stream=None) -> None:
        """``send_g [cp, rows(me), C]``: chunk d (d != me) goes to rank d's ``recv[g, rows_before(me):, :]``."""
        cp, me = self.size, self.rank
        assert tuple(send_g.shape) == (cp, self._rows[me], self._C)
        order = self._order()
        if not order:
            return
        off = self._recv_off + (g * self._L + self._row_off[me]) * self._C * self._item
        self.window.push([(d, off, send_g[d]) for d in order],
                         [(d, self.flag_index(0, g, me, cp)) for d in order], self.epoch, stream)

    def push_in_local(self, send: torch.Tensor, stream=None) -> None:
        """``send [G, cp, rows(me), C]``: this rank's own chunk of every group into its own ``recv``."""
        cp, me, G = self.size, self.rank, self._G
        assert tuple(send.shape) == (G, cp, self._rows[me], self._C)
        copies = [(me, self._recv_off + (g * self._L + self._row_off[me]) * self._C * self._item, send[g, me])
                  for g in range(G)]
        self.window.push(copies, [(me, self.flag_index(0, g, me, cp)) for g in range(G)], self.epoch, stream)

    def wait_in(self, g_first: int, g_last: int, stream=None) -> None:
        cp = self.size
        self.window.wait(self.flag_index(0, g_first, 0, cp), (g_last - g_first + 1) * cp, self.epoch, stream)

    def push_out(self, g: int, o: torch.Tensor, stream=None) -> None:
        """``o [L, w]`` (all tokens, my heads of group g): rows of rank d (d != me) go to rank d's ``back[g, me]``."""
        cp, me = self.size, self.rank
        assert tuple(o.shape) == (self._L, self._w) and o.is_contiguous()
        order = self._order()
        if not order:
            return
        copies = []
        for d in order:
            off = self._back_off + (g * cp + me) * self._rows[d] * self._w * self._item
            copies.append((d, off, o[self._row_off[d]:self._row_off[d] + self._rows[d]]))
        self.window.push(copies, [(d, self.flag_index(1, g, me, cp)) for d in order], self.epoch, stream)

    def push_out_local(self, outs: Sequence[Tuple[int, torch.Tensor]], stream=None) -> None:
        """``outs``: (group g, ``o [L, w]``) for every group: this rank's own rows into its own ``back[g, me]``."""
        cp, me = self.size, self.rank
        lo, n = self._row_off[me], self._rows[me]
        copies, flags = [], []
        for g, o in outs:
            assert tuple(o.shape) == (self._L, self._w) and o.is_contiguous()
            copies.append((me, self._back_off + (g * cp + me) * n * self._w * self._item, o[lo:lo + n]))
            flags.append((me, self.flag_index(1, g, me, cp)))
        self.window.push(copies, flags, self.epoch, stream)

    def wait_out(self, stream=None) -> None:
        cp = self.size
        self.window.wait(self.flag_index(1, 0, 0, cp), self._G * cp, self.epoch, stream)


_EXCHANGES = {}  # one set of windows per (process group, device) for the life of the process


def make_cuda_exchange(group, rank: int, size: int, device: torch.device) -> Optional[PeerExchange]:
    """The ``PeerExchange`` over CUDA IPC windows for ``group`` on ``device``, or ``None`` (with a warning, on every
    rank alike) when the windows cannot be mapped -- the caller then keeps the NCCL all-to-all.  First call per group
    is collective."""
    key = (getattr(group, "group_name", None) or id(group), str(device), rank, size)
    if key not in _EXCHANGES:
        px = PeerExchange(CudaIpcWindow(group, rank, size, device), rank, size)
        try:
            px.window.ensure(FLAG_BYTES + (1 << 20))  # collective probe: allocate, exchange handles, map every peer
        except PeerUnavailable as e:
            warnings.warn(f"dualforce_b200: peer-memory exchange unavailable, using NCCL all-to-all ({e})")
            px = None
        _EXCHANGES[key] = px
    return _EXCHANGES[key]
