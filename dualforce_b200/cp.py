"""Context parallelism (``cp_size`` GPUs of one NVSwitch box) for the dual-tower forward.

Reference behaviour (mova/distributed/functional.py:55-111, pipeline_mova.py:653-674,704-706, and yunchang's
LongContextAttention behind USPAttention, wan_video_dit.py:192-208): chunk + zero-pad BOTH token sequences, run
every attention as Ulysses(<=4) x ring, all-gather at the end.  The padded audio tokens are not masked, so the
reference's cp>1 result differs from its own cp=1 result.

This design targets the cp=1 result exactly:
  * video tokens are sharded in contiguous, possibly ragged chunks (no padding, nothing to mask),
  * video self-attention is Ulysses over ALL cp ranks (40 heads -> 20/10/5 per rank): one all-to-all of the fused
    q|k|v buffer in, one all-to-all of the attention output back, both in place thanks to the segmented GEMM
    operands (``ops.linear(..., out_segments=)`` writes destination-rank-major, ``ops.linear(..., segments=)`` reads
    source-rank-major), optionally split in head groups so group g+1 travels while group g is in the tensor cores,
  * the 403-token audio tower, the text keys and the a2v keys are replicated: no communication,
  * v2a (audio queries x sharded video keys): every rank attends over its own keys, then an all-gather of
    (partial output, log-sum-exp) -- 403 x 12 x 129 values -- and an exact LSE merge,
  * one all-gather of the final video hidden states.

This file holds the layout / index logic and the collectives (pure torch.distributed, testable on CPU with gloo);
the kernels that consume those layouts are called from :mod:`dualforce_b200.pipeline`.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def seq_chunks(total: int, cp: int) -> List[Tuple[int, int]]:
    """Contiguous [start, stop) token ranges per rank; same chunk length rule as torch.chunk in
    _sp_split_tensor (functional.py:55-58) but the last chunk is simply shorter instead of zero-padded."""
    chunk = -(-total // cp)
    out = []
    for r in range(cp):
        s = min(r * chunk, total)
        out.append((s, min(s + chunk, total)))
    if any(b <= a for a, b in out):
        raise ValueError(f"context parallel size {cp} leaves a rank without tokens for a sequence of {total}")
    return out


class UlyssesPlan:
    """Head <-> sequence redistribution plan for one attention of ``num_heads`` heads over ``cp`` ranks, optionally
    pipelined in ``groups`` head groups.

    Segment order everywhere is ``[group g][rank r][local head j][128 channels]``, i.e. head
    ``r * Hc + g * Hg + j`` with ``Hc = H / cp`` heads per rank and ``Hg = Hc / groups`` heads per group.
    """

    def __init__(self, num_heads: int, head_dim: int, cp: int, groups: int = 1):
        if num_heads % cp:
            raise ValueError(f"{num_heads} heads cannot be split over {cp} ranks")
        self.H, self.D, self.cp = num_heads, head_dim, cp
        self.Hc = num_heads // cp
        if self.Hc % groups:
            raise ValueError(f"{self.Hc} heads per rank cannot be split in {groups} groups")
        self.groups = groups
        self.Hg = self.Hc // groups
        self.w = self.Hg * head_dim  # channels of one (group, rank) segment
        self.nseg = groups * cp

    @staticmethod
    def pick_groups(heads_per_rank: int, want: int = 2) -> int:
        """Largest divisor of heads_per_rank that is <= want (5 heads -> 1, 10 -> 2, 20 -> 2)."""
        for g in range(min(want, heads_per_rank), 0, -1):
            if heads_per_rank % g == 0:
                return g
        return 1

    def head_order(self) -> torch.Tensor:
        idx = []
        for g in range(self.groups):
            for r in range(self.cp):
                for j in range(self.Hg):
                    idx.append(r * self.Hc + g * self.Hg + j)
        return torch.tensor(idx, dtype=torch.long)

    def channel_index(self) -> torch.Tensor:
        """Permutation of the d = H*D channels into segment order (norm weights, o-proj input columns)."""
        heads = self.head_order()
        return (heads[:, None] * self.D + torch.arange(self.D)[None, :]).reshape(-1)

    def qkv_row_index(self) -> torch.Tensor:
        """Rows of ``cat(Wq, Wk, Wv)`` ([3d, d]) in send order ``[g][r][q|k|v][j][128]``."""
        d = self.H * self.D
        idx = []
        for g in range(self.groups):
            for r in range(self.cp):
                for part in range(3):
                    for j in range(self.Hg):
                        h = r * self.Hc + g * self.Hg + j
                        idx.append(part * d + h * self.D + torch.arange(self.D))
        return torch.cat(idx)


def head_group_sets(heads_per_rank: int, want: int = 2, set_sizes: Optional[Sequence[int]] = None):
    """(number of head groups, attention sets) for the pipelined Ulysses exchange of ``heads_per_rank`` heads.

    The exchange buffers are laid out in equal head groups (a constraint of the segmented GEMM operands).  When the
    heads split evenly in ``want`` groups, every group is one exchange and one attention launch (20 heads -> 2 x 10,
    10 -> 2 x 5: the round-1 configuration).  An odd count (cp = 8: 5 heads per rank) is exchanged head by head and
    attended in SETS of consecutive heads, one attention launch per set (the heads are its batch dimension), the
    launches alternating between two side streams so the partial last wave of one launch (337 query tiles on 148 SMs)
    overlaps the first wave of the next.  ``set_sizes`` partitions the heads (``(1, 3, 1)`` -> ``[[0], [1, 2, 3], [4]]``:
    a short first set so attention starts after one head has landed, a short last set so little of the return exchange
    is exposed); ``None`` = one set per head, the configuration measured with the NCCL exchange
    (profiles/r02_timeline_cp8.json)."""
    g = UlyssesPlan.pick_groups(heads_per_rank, want)
    if g >= min(want, heads_per_rank) or heads_per_rank < 3:
        return g, [[i] for i in range(g)]
    if set_sizes is not None:
        if sum(set_sizes) != heads_per_rank or min(set_sizes) < 1:
            raise ValueError(f"attention set sizes {tuple(set_sizes)} do not partition {heads_per_rank} heads")
        sets, i = [], 0
        for n in set_sizes:
            sets.append(list(range(i, i + n)))
            i += n
        return heads_per_rank, sets
    return heads_per_rank, [[i] for i in range(heads_per_rank)]


def all_to_all_rows(inp: torch.Tensor, in_rows: Sequence[int], out_rows: Sequence[int],
                    group: Optional[dist.ProcessGroup], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """all_to_all_single over dim 0 of a ``[sum(in_rows), C]`` matrix: rows ``in_rows[r]`` go to rank r, the result
    is ``[sum(out_rows), C]`` with rank r's rows at offset ``sum(out_rows[:r])``.  Pass ``out`` when the call is
    queued on a side stream, so the buffer belongs to the consumer's stream in the caching allocator."""
    if out is None:
        out = torch.empty((sum(out_rows),) + tuple(inp.shape[1:]), dtype=inp.dtype, device=inp.device)
    assert out.is_contiguous() and out.shape[0] == sum(out_rows)
    dist.all_to_all_single(out, inp, output_split_sizes=list(out_rows), input_split_sizes=list(in_rows), group=group)
    return out


def scatter_heads(send: torch.Tensor, rows_per_rank: Sequence[int], rank: int,
                  group: Optional[dist.ProcessGroup], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Forward Ulysses exchange of one head group.  ``send``: ``[cp, Lc, C]`` destination-rank-major (C = 3*w for the
    fused q|k|v buffer); returns ``[L, C]``: all tokens, this rank's heads of the group."""
    cp, Lc, C = send.shape
    assert Lc == rows_per_rank[rank]
    return all_to_all_rows(send.reshape(cp * Lc, C), [Lc] * cp, list(rows_per_rank), group, out=out)


def gather_heads(o: torch.Tensor, rows_per_rank: Sequence[int], rank: int,
                 group: Optional[dist.ProcessGroup], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Inverse exchange.  ``o``: ``[L, w]`` (all tokens, this rank's heads); returns ``[cp, Lc, w]``
    source-rank-major: this rank's tokens, every rank's heads of the group."""
    cp = len(rows_per_rank)
    Lc = rows_per_rank[rank]
    if out is not None:
        out = out.view(cp * Lc, o.shape[-1])
    out = all_to_all_rows(o, list(rows_per_rank), [Lc] * cp, group, out=out)
    return out.view(cp, Lc, o.shape[-1])


def all_gather_cat(x: torch.Tensor, rows_per_rank: Sequence[int], group: Optional[dist.ProcessGroup],
                   dim: int = 1) -> torch.Tensor:
    """Concatenate ragged per-rank chunks along ``dim`` (the final gather of pipeline_mova.py:704-706, without the
    pad-and-strip of functional.py:106-111)."""
    cp = len(rows_per_rank)
    rmax = max(rows_per_rank)
    xt = x.transpose(0, dim).contiguous()
    if xt.shape[0] < rmax:
        pad = torch.zeros((rmax - xt.shape[0],) + tuple(xt.shape[1:]), dtype=x.dtype, device=x.device)
        xt = torch.cat([xt, pad], dim=0)
    buf = torch.empty((cp * rmax,) + tuple(xt.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(buf, xt, group=group)
    if all(r == rmax for r in rows_per_rank):
        full = buf
    else:
        full = torch.cat([buf[r * rmax:r * rmax + rows_per_rank[r]] for r in range(cp)], dim=0)
    return full.transpose(0, dim)


def all_gather_stack(x: torch.Tensor, cp: int, group: Optional[dist.ProcessGroup]) -> torch.Tensor:
    """``[cp, *x.shape]`` stack of every rank's ``x`` (partial outputs / LSEs of the sharded v2a attention)."""
    x = x.contiguous()
    flat = torch.empty((cp * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(flat, x, group=group)  # concatenated along dim 0 (the form every backend accepts)
    return flat.view((cp,) + tuple(x.shape))


def ulysses_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int, attn_fn,
                      group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Generic head <-> sequence redistribution around a local attention call, for callers that hold sequence shards
    ``[B, S/P, H*D]`` and want the reference's ``USPAttention.forward`` contract (wan_video_dit.py:203-208: yunchang
    LongContextAttention = all-to-all in, attention over H/P heads and the full sequence, all-to-all out).  Equal
    shard lengths on every rank are required here (the zero-copy path in pipeline.py handles ragged video chunks).
    ``attn_fn(q, k, v, heads)`` works on flat ``[B, S, heads*D]`` tensors."""
    P = dist.get_world_size(group) if dist.is_initialized() else 1
    if P == 1:
        return attn_fn(q, k, v, num_heads)
    B, Sl, HD = q.shape
    if num_heads % P:
        raise ValueError(f"{num_heads} heads cannot be split over {P} ranks")
    D = HD // num_heads
    Hc = num_heads // P

    def scatter(t: torch.Tensor) -> torch.Tensor:
        # [B, Sl, P, Hc*D] -> [P, B, Sl, Hc*D] (destination-rank-major) -> exchange -> [P(src), B, Sl, Hc*D]
        send = t.reshape(B, t.shape[1], P, Hc * D).permute(2, 0, 1, 3).contiguous()
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
        # source rank r holds tokens [r*Sl, (r+1)*Sl): concatenate along the sequence
        return recv.permute(1, 0, 2, 3).reshape(B, P * t.shape[1], Hc * D)

    o = attn_fn(scatter(q), scatter(k), scatter(v), Hc)  # [B, S, Hc*D]
    send = o.reshape(B, P, Sl, Hc * D).permute(1, 0, 2, 3).contiguous()  # [P(dst: its tokens), B, Sl, Hc*D]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    # recv[src] = this rank's tokens, heads of rank src -> [B, Sl, P*Hc*D]
    return recv.permute(1, 2, 0, 3).reshape(B, Sl, HD)


# ----------------------------------------------------------------------------------------------------------------
# Forward-only twins of the reference's sequence-parallel helpers (mova/distributed/functional.py:55-121), for callers
# that keep the reference loop (pad-and-strip semantics, same return tuples).  Pure index / byte plumbing: bit-exact
# with the reference.  ``forward_dual_tower_dit`` itself does not use them (it shards ragged, never pads -- see the
# module docstring); the autograd halves (_AllGather / _AllGatherAvg.backward) belong to training and are not built.
# ----------------------------------------------------------------------------------------------------------------
def _sp_split_tensor(x: torch.Tensor, *, sp_size: int, sp_rank: int, dim: int = 1):
    """functional.py:55-73: ``torch.chunk`` along ``dim``, zero chunk for surplus ranks, zero-pad a short chunk.
    Returns ``(chunk, chunk_len, pad_len, total_len)``."""
    total_len = x.shape[dim]
    chunks = torch.chunk(x, sp_size, dim=dim)
    chunk_len = chunks[0].shape[dim]
    if sp_rank < len(chunks):
        chunk = chunks[sp_rank]
    else:
        shape = list(x.shape)
        shape[dim] = chunk_len
        chunk = x.new_zeros(shape)
    if chunk.shape[dim] < chunk_len:
        shape = list(chunk.shape)
        shape[dim] = chunk_len - chunk.shape[dim]
        chunk = torch.cat([chunk, x.new_zeros(shape)], dim=dim)
    return chunk, chunk_len, chunk_len * sp_size - total_len, total_len


def _sp_split_tensor_dim_0(x: torch.Tensor, *, sp_size: int, sp_rank: int):
    """functional.py:76-96 (the RoPE tables are split along dim 0)."""
    return _sp_split_tensor(x, sp_size=sp_size, sp_rank=sp_rank, dim=0)


def _sp_all_gather(x_local: torch.Tensor, *, sp_group: Optional[dist.ProcessGroup], pad_len: int) -> torch.Tensor:
    """functional.py:99-104 forward: all-gather the equal-length chunks, concatenate along dim 1, strip the padding."""
    parts = [torch.empty_like(x_local) for _ in range(dist.get_world_size(group=sp_group))]
    dist.all_gather(parts, x_local.contiguous(), group=sp_group)
    gathered = torch.cat(parts, dim=1)
    return gathered[:, :-pad_len] if pad_len > 0 else gathered


def _sp_all_gather_avg(x_local: torch.Tensor, *, sp_group: Optional[dist.ProcessGroup], pad_len: int) -> torch.Tensor:
    """functional.py:107-112 forward (the AVG only concerns the backward reduce-scatter): same as ``_sp_all_gather``."""
    return _sp_all_gather(x_local, sp_group=sp_group, pad_len=pad_len)


def _sp_select_rank(x_global: torch.Tensor, *, sp_size: int, sp_rank: int, chunk_len: int, pad_len: int):
    """functional.py:115-121: this rank's (padded) slice of a full-length tensor."""
    total_padded = chunk_len * sp_size
    if pad_len > 0 and x_global.shape[1] != total_padded:
        x_global = torch.nn.functional.pad(x_global, (0, 0, 0, total_padded - x_global.shape[1]))
    start = sp_rank * chunk_len
    return x_global[:, start:start + chunk_len]
