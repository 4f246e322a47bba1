"""Drop-in for ``MOVA.forward_dual_tower_dit`` (mova/diffusion/pipelines/pipeline_mova.py:612-711) and the
installer that swaps the B200 modules into an existing ``MOVA`` pipeline the way the reference's own
``MOVA.replace_attention`` (pipeline_mova.py:124-148) swaps attention modules.

``forward_dual_tower_dit`` keeps the reference signature.  With ``cp_mesh=None`` it is the reference loop
(bridge -> video block -> audio block for the first min(layers) layers, then the remaining video blocks) on the
fused kernels.  With a ``cp_mesh`` it runs the context-parallel design of :mod:`dualforce_b200.cp`: sharded video
tokens, Ulysses all-to-all overlapped with the attention kernels on a side stream, replicated audio tower, LSE
merge for the v2a bridge -- and returns the same full-length tensors as cp=1.
"""
from __future__ import annotations

import contextlib
import os
import types
from typing import Optional, Sequence, Tuple

import torch

from . import cp as cpmod
from . import ops, rope
from .modules import ConditionalCrossAttentionBlock, DiTBlock, DualTowerConditionalBridge, param_sig, prepack

__all__ = ["forward_dual_tower_dit", "install", "swap_modules", "CPRuntime", "GraphedForward"]


# ----------------------------------------------------------------------------------------------------------------
# context-parallel runtime
# ----------------------------------------------------------------------------------------------------------------
class CPRuntime:
    """Process-group handle + the communication stream the all-to-alls are queued on."""

    audio_side_stream = True  # class-level switch for A/B measurements and bisecting: False = everything on one stream
    # Data path of the Ulysses exchange: "peer" = copy-engine pushes into the peers' windows + flag words
    # (dualforce_b200/peer.py; falls back to NCCL with a warning when the windows cannot be mapped), "nccl" =
    # dist.all_to_all_single on the communication stream (round-1 path, kept for A/B measurements and CUDA graphs).
    exchange = "peer"
    # Attention sets when the heads per rank do not split in two groups (cp = 8: 5 heads), see cp.head_group_sets;
    # None = (1, n - 2, 1) with the peer exchange, one set per head with NCCL.
    set_sizes = None
    # Streams the remote pushes of the peer exchange are dealt over (by destination): a copy-engine operation costs
    # ~9 us on its stream and 14 of them make one head group's exchange at cp = 8.  0 = one stream per peer at cp >= 8 (measured
    # at cp = 8, profiles/r02_timeline_cp8_push_streams.json: 300.5 ms per forward against 302.1-304.5 with one stream,
    # bench 1.655 against 1.623 steps/s).
    push_streams_n = 0

    def __init__(self, group, rank: int, size: int, device: torch.device, head_groups: Optional[int] = None):
        self.group, self.rank, self.size, self.device = group, rank, size, device
        # head groups the Ulysses exchange is pipelined in when the heads per rank split evenly (2: measured in round 1
        # at cp = 2 / 4); an odd count (5 heads per rank at cp = 8) is exchanged head by head and attended in two sets
        self.head_groups = 2 if head_groups is None else head_groups
        cuda = device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=device) if cuda else None
        # the replicated audio tower and the v2a bridge direction run here, beside the video block of the same layer
        self.audio_stream = torch.cuda.Stream(device=device) if (cuda and self.audio_side_stream) else None
        # more than two attention sets per layer (odd head count per rank): launches alternate between these
        self.attn_streams = [torch.cuda.Stream(device=device) for _ in range(2)] if cuda else []
        self._px = None  # PeerExchange, False once it proved unavailable (tests inject a shared-memory one)
        self._push_streams = [self.comm_stream] if cuda else []

    def push_streams(self):
        """The communication stream plus ``push_streams_n - 1`` more (created on first use, capped at 8)."""
        # 0: one stream per peer where the exchange is latency-bound (many small chunks: measured at cp = 8); with 2 or
        # 4 ranks the chunks are 60-165 MB each, a single stream already runs at NVLink bandwidth (measured at cp = 2)
        want = int(self.push_streams_n) or (self.size - 1 if self.size >= 8 else 1)
        n = max(1, min(want, 8, max(self.size - 1, 1)))
        while self.device.type == "cuda" and len(self._push_streams) < n:
            self._push_streams.append(torch.cuda.Stream(device=self.device))
        return self._push_streams[:n]

    def peer_exchange(self):
        """The peer-memory exchange for this runtime, or None (NCCL path).  First use is collective."""
        if self._px is None:
            self._px = False
            if self.exchange == "peer" and self.device.type == "cuda":
                from . import peer

                self._px = peer.make_cuda_exchange(self.group, self.rank, self.size, self.device) or False
        if self._px is False or self.exchange != "peer":
            return None
        if self.device.type == "cuda" and torch.cuda.is_current_stream_capturing():
            return None  # epochs and window growth are host-side state: a captured graph keeps the NCCL exchange
        return self._px

    def attention_sets(self, heads_per_rank: int, peer: bool):
        sizes = self.set_sizes
        if sizes is None and peer and heads_per_rank >= 3:
            sizes = (1, heads_per_rank - 2, 1)
        return cpmod.head_group_sets(heads_per_rank, self.head_groups, sizes)

    @classmethod
    def from_mesh(cls, cp_mesh, device: torch.device, head_groups: Optional[int] = None) -> "CPRuntime":
        key = (id(cp_mesh), str(device), head_groups)
        rt = _RUNTIMES.get(key)
        if rt is None:
            rt = cls(cp_mesh.get_group(), cp_mesh.get_local_rank(), cp_mesh.size(), device, head_groups)
            _RUNTIMES[key] = rt
        return rt


_RUNTIMES = {}

# Optional per-segment device timeline (benchmarks/cp_layer_timeline.py): a list that receives
# (segment name, stream name, start event, end event); None = off (the normal state: no events are created).
TIMELINE = None


@contextlib.contextmanager
def _seg(name: str, stream_name: str = "main"):
    if TIMELINE is None:
        yield
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    try:
        yield
    finally:
        e1.record()
        TIMELINE.append((name, stream_name, e0, e1))


def _cp_weights(block: DiTBlock, plan: cpmod.UlyssesPlan):
    """Destination-rank-major copies of the self-attention weights of ``block`` for ``plan``: built once per
    (plan, parameter pointers AND versions) -- ``load_state_dict`` / ``copy_`` / an in-place LoRA merge after the first
    context-parallel forward bump the versions and force a rebuild.  Cost: a second copy of Wq|Wk|Wv|Wo per video block
    (4 d^2 bf16 = 210 MB at d = 5120, 8.4 GB per 40-layer tower) while cp > 1; one entry per block, replaced when cp
    changes."""
    cache = getattr(block, "_cp_cache", None)
    if cache is None:
        cache = block._cp_cache = {}
    sa = block.self_attn
    key = (plan.cp, plan.groups) + param_sig(sa.q.weight, sa.k.weight, sa.v.weight, sa.o.weight, sa.q.bias, sa.k.bias,
                                             sa.v.bias, sa.norm_q.weight, sa.norm_k.weight)
    hit = cache.get(key)
    if hit is not None:
        return hit
    dev = sa.q.weight.device
    rows = plan.qkv_row_index().to(dev)
    chan = plan.channel_index().to(dev)
    w = torch.cat([sa.q.weight.data, sa.k.weight.data, sa.v.weight.data], dim=0).index_select(0, rows).contiguous()
    b = torch.cat([sa.q.bias.data, sa.k.bias.data, sa.v.bias.data], dim=0).index_select(0, rows).contiguous()
    nq = sa.norm_q.weight.data.index_select(0, chan).contiguous()
    nk = sa.norm_k.weight.data.index_select(0, chan).contiguous()
    wo = sa.o.weight.data.index_select(1, chan).contiguous()
    cache.clear()
    cache[key] = (w, b, nq, nk, wo)
    return cache[key]


def _self_attention_cp(block: DiTBlock, h: torch.Tensor, tables, x: torch.Tensor, gate: torch.Tensor,
                       rt: CPRuntime, rows_per_rank: Sequence[int]) -> torch.Tensor:
    """``x + gate * self_attn(h)`` for the local token chunk: QKV GEMM writing destination-rank-major, full-row
    RMSNorm + RoPE on the segmented buffer, per head group {all-to-all, attention, all-to-all back} with the
    exchanges on the communication stream, o-projection reading source-rank-major with the gated residual fused."""
    sa = block.self_attn
    px = rt.peer_exchange()
    groups, sets = rt.attention_sets(sa.num_heads // rt.size, px is not None)
    plan = cpmod.UlyssesPlan(sa.num_heads, sa.head_dim, rt.size, groups)
    w, b, nq, nk, wo = _cp_weights(block, plan)
    Lc = h.shape[1]
    G, cp, wd = plan.groups, plan.cp, plan.w
    cos, sin = tables
    with _seg("qkv_gemm+qknorm_rope"):
        send = ops.linear(h[0], w, b, out_segments=plan.nseg)  # [G*cp, Lc, 3*wd]
        seg_stride = Lc * 3 * wd
        ops.rmsnorm_rope_(send[0][:, 0:wd], nq, sa.norm_q.eps, head_dim=sa.head_dim, cos=cos, sin=sin,
                          rope_mode=ops.ROPE_INTERLEAVED, segments=plan.nseg, seg_stride=seg_stride)
        ops.rmsnorm_rope_(send[0][:, wd:2 * wd], nk, sa.norm_k.eps, head_dim=sa.head_dim, cos=cos, sin=sin,
                          rope_mode=ops.ROPE_INTERLEAVED, segments=plan.nseg, seg_stride=seg_stride)
    send = send.view(G, cp, Lc, 3 * wd)
    comm = rt.comm_stream  # None for CPU tensors (gloo host-logic tests): same exchanges, program order
    main = torch.cuda.current_stream() if comm is not None else None

    def on_comm():
        return torch.cuda.stream(comm) if comm is not None else contextlib.nullcontext()

    def record(stream):
        if comm is None:
            return None
        ev = torch.cuda.Event()
        ev.record(stream)
        return ev

    def wait(stream, ev):
        if ev is not None:
            stream.wait_event(ev)

    ready = record(main)
    L = sum(rows_per_rank)
    if px is not None:
        # this rank's window: the peers write recv / back directly, flag words say when (dualforce_b200/peer.py)
        recv, back = px.begin(G, L, 3 * wd, rows_per_rank, wd)
    else:
        # every buffer is allocated on the compute stream; the side stream only fills it between two events
        recv = torch.empty(G, L, 3 * wd, dtype=torch.bfloat16, device=h.device)
        back = torch.empty(G, cp, Lc, wd, dtype=torch.bfloat16, device=h.device)
    outs = [torch.empty(len(gs), L, wd, dtype=torch.bfloat16, device=h.device) for gs in sets]
    in_done = []
    pushers = rt.push_streams() if (px is not None and comm is not None) else None  # pushers[0] is `comm`
    with on_comm():
        wait(comm, ready)
        if px is not None:
            for ps in (pushers or [])[1:]:
                wait(ps, ready)
            px.push_in_local(send, stream=comm)  # same-device copies first: nothing on the SMs is in their way yet
        for g in range(G):
            with _seg(f"all_to_all_in[{g}]", "comm"):
                if px is not None:
                    px.push_in(g, send[g], stream=pushers if pushers else comm)
                else:
                    cpmod.scatter_heads(send[g], rows_per_rank, rt.rank, rt.group, out=recv[g])  # [L, 3*wd]
            in_done.append(record(comm))
    out_done = []
    # Side streams for the attention launches: with more than two sets, or with a set of one or two heads (337 query
    # tiles per head on 148 SMs leave a partial last wave that only a launch on ANOTHER stream can fill).
    few = min(len(gs) * plan.Hg for gs in sets) <= 2
    side = rt.attn_streams if (comm is not None and (len(sets) > 2 or (few and len(sets) > 1))) else []
    for idx, (gs, o) in enumerate(zip(sets, outs)):
        # one attention launch per SET of consecutive head groups: the groups are its batch dimension (stride L*3*wd),
        # so a set starts as soon as its last group has landed while the next set is still on the wire.
        st = side[idx % len(side)] if side else main
        qkv = recv[gs[0]:gs[-1] + 1]
        with (torch.cuda.stream(st) if side else contextlib.nullcontext()):
            with _seg(f"wait_all_to_all_in{gs}", "attn" if side else "main"):
                if px is not None:
                    if side:
                        wait(st, ready)  # stream order behind this layer's QKV GEMM (and the previous layer)
                    px.wait_in(gs[0], gs[-1], stream=st)
                else:
                    wait(st, in_done[gs[-1]])
            with _seg(f"self_attention{gs}", "attn" if side else "main"):
                ops.attention(qkv[..., 0:wd], qkv[..., wd:2 * wd], qkv[..., 2 * wd:], plan.Hg, out=o)
            att = record(st)
        with on_comm():
            wait(comm, att)
            for ps in (pushers or [])[1:]:
                wait(ps, att)
            with _seg(f"all_to_all_out{gs}", "comm"):
                for i, g in enumerate(gs):
                    if px is not None:
                        px.push_out(g, o[i], stream=pushers if pushers else comm)
                    else:
                        cpmod.gather_heads(o[i], rows_per_rank, rt.rank, rt.group, out=back[g])
            out_done.append(record(comm))
    for ps in (pushers or [])[1:]:  # the other push streams: one event after their last push covers all of them
        out_done.append(record(ps))
    with _seg("wait_all_to_all_out"):
        for ev in out_done:
            wait(main, ev)
        if px is not None:
            px.push_out_local([(g, o[i]) for gs, o in zip(sets, outs) for i, g in enumerate(gs)], stream=main)
            px.wait_out(stream=main)
    with _seg("o_proj"):
        return ops.linear(back.view(G * cp, Lc, wd), wo, sa.o.bias, epilogue=ops.EPI_RESIDUAL, residual=x[0], gate=gate,
                          segments=plan.nseg).unsqueeze(0)


def _video_block_cp(block: DiTBlock, x: torch.Tensor, context: torch.Tensor, t_mod: torch.Tensor, tables,
                    rt: CPRuntime, rows_per_rank: Sequence[int]) -> torch.Tensor:
    """DiTBlock.forward (wan_video_dit.py:275-291) on this rank's token chunk; only the self-attention communicates."""
    with _seg("modulation+ln1"):
        mod = block.modulation_f32(t_mod)
        ca = block.cross_attn
        h = ops.layernorm(x, block.norm1.eps, shift=mod[0], scale=mod[1])
    x = _self_attention_cp(block, h, tables, x, mod[2], rt, rows_per_rank)
    with _seg("text_cross_attention"):
        h = ops.layernorm(x, block.norm3.eps, weight=block.norm3.weight, bias=block.norm3.bias, out=h)
        a = ca.attend(h, context)
        ops.linear(a, ca.o.weight, ca.o.bias, epilogue=ops.EPI_RESIDUAL, residual=x, out=x)
    with _seg("ffn"):
        ops.layernorm(x, block.norm2.eps, shift=mod[3], scale=mod[4], out=h)
        u = ops.linear(h, block.ffn[0].weight, block.ffn[0].bias, epilogue=ops.EPI_GELU_TANH)
        ops.linear(u, block.ffn[2].weight, block.ffn[2].bias, epilogue=ops.EPI_RESIDUAL, residual=x, gate=mod[5], out=x)
    return x


def _v2a_cp(cond: ConditionalCrossAttentionBlock, audio_x: torch.Tensor, x_loc: torch.Tensor, a_tables, v_tables_loc,
            scale: float, rt: CPRuntime) -> torch.Tensor:
    """v2a bridge direction with sharded video keys: local partial attention, all-gather of (o, lse), exact merge."""
    inner = cond.inner
    q = inner.project_q(audio_x, a_tables)
    k, v = inner.project_kv(cond.normed_condition(x_loc), v_tables_loc)
    o, lse = ops.attention(q, k, v, inner.num_heads, return_lse=True)
    o_all = cpmod.all_gather_stack(o[0], rt.size, rt.group)      # [cp, L_a, d_a]
    lse_all = cpmod.all_gather_stack(lse[0], rt.size, rt.group)  # [cp, H, L_a]
    merged = ops.lse_merge(o_all, lse_all, inner.num_heads)
    return ops.linear(merged.unsqueeze(0), inner.o.weight, inner.o.bias, epilogue=ops.EPI_RESIDUAL, residual=audio_x,
                      scale=float(scale))


# ----------------------------------------------------------------------------------------------------------------
# the path
# ----------------------------------------------------------------------------------------------------------------
def _check_modules(visual_dit, audio_dit, bridge) -> None:
    for blk in list(visual_dit.blocks) + list(audio_dit.blocks):
        if not isinstance(blk, DiTBlock):
            raise TypeError("forward_dual_tower_dit needs dualforce_b200.DiTBlock modules; call dualforce_b200.install(pipe) "
                            f"first (found {type(blk).__module__}.{type(blk).__name__})")
    if not isinstance(bridge, DualTowerConditionalBridge):
        raise TypeError("forward_dual_tower_dit needs a dualforce_b200.DualTowerConditionalBridge; call install(pipe)")


@torch.no_grad()
def _forward_eager(self, visual_dit, visual_x: torch.Tensor, audio_x: torch.Tensor,
                           visual_context: torch.Tensor, audio_context: torch.Tensor, visual_t_mod: torch.Tensor,
                           audio_t_mod: Optional[torch.Tensor], visual_freqs: torch.Tensor, audio_freqs: torch.Tensor,
                           grid_size: Tuple[int, int, int], video_fps: float, condition_scale: Optional[float] = 1.0,
                           a2v_condition_scale: Optional[float] = None, v2a_condition_scale: Optional[float] = None,
                           cp_mesh=None, _gather: bool = True):
    """Same contract as pipeline_mova.py:612-711: returns full-length ``(visual_x, audio_x)`` hidden states.
    ``self`` only needs ``audio_dit`` and ``dual_tower_bridge`` attributes (a ``MOVA`` pipeline after ``install``).

    ``_gather=False`` (context parallel only, used by ``step.inference_single_step``) skips the final all-gather and
    returns ``(visual_x_local, audio_x, rows_per_rank, group)`` so the head can run on the local chunk."""
    audio_dit, bridge = self.audio_dit, self.dual_tower_bridge
    _check_modules(visual_dit, audio_dit, bridge)
    if visual_x.shape[0] != 1 and cp_mesh is not None:
        raise NotImplementedError("context parallelism with batch > 1 (MOVA runs CFG as two B=1 forwards)")
    min_layers = min(len(visual_dit.blocks), len(audio_dit.blocks))
    visual_layers = len(visual_dit.blocks)
    assert visual_layers >= min_layers, "visual_layers must be greater than min_layers"

    v_tab = rope.as_tables(visual_freqs)
    a_tab = rope.as_tables(audio_freqs)
    v_cs = a_cs = None
    if bridge.apply_cross_rope:
        # built (and memoised) in fp32: the reference rounds these tables to bf16 (interactionv2.py:236-237),
        # which only adds noise relative to the fp32 oracle
        v_pair, a_pair = bridge.build_aligned_freqs(video_fps=video_fps, grid_size=grid_size,
                                                    audio_steps=audio_x.shape[1], device=visual_x.device,
                                                    dtype=torch.float32)
        v_cs, a_cs = rope.as_tables(v_pair), rope.as_tables(a_pair)

    def scale_for(direct):
        return bridge._scale(direct if direct is not None else condition_scale)

    if cp_mesh is None:
        for i in range(min_layers):
            if bridge.should_interact(i, "a2v"):
                visual_x, audio_x = bridge(i, visual_x, audio_x, x_freqs=v_cs, y_freqs=a_cs,
                                           a2v_condition_scale=a2v_condition_scale,
                                           v2a_condition_scale=v2a_condition_scale, condition_scale=condition_scale,
                                           video_grid_size=grid_size)
            visual_x = visual_dit.blocks[i](visual_x, visual_context, visual_t_mod, v_tab)
            audio_x = audio_dit.blocks[i](audio_x, audio_context, audio_t_mod, a_tab)
        for i in range(min_layers, visual_layers):
            visual_x = visual_dit.blocks[i](visual_x, visual_context, visual_t_mod, v_tab)
        return visual_x, audio_x

    # ------------------------------ context parallel ------------------------------
    rt = CPRuntime.from_mesh(cp_mesh, visual_x.device)
    chunks = cpmod.seq_chunks(visual_x.shape[1], rt.size)
    rows = [b - a for a, b in chunks]
    s0, s1 = chunks[rt.rank]
    x_loc = visual_x[:, s0:s1].contiguous()
    v_tab_loc = (v_tab[0][s0:s1].contiguous(), v_tab[1][s0:s1].contiguous())
    v_cs_loc = (v_cs[0][s0:s1].contiguous(), v_cs[1][s0:s1].contiguous()) if v_cs is not None else None
    # Two streams per layer.  The video side (a2v bridge direction + video block, with its all-to-alls) stays on the
    # current stream; the replicated audio side (v2a bridge direction with its small all-gathers + audio block: ~35
    # launch-bound kernels on 403 tokens that do not shrink with cp) runs on rt.audio_stream beside it.  Both bridge
    # directions read the PRE-bridge states of the other tower, so each layer starts with one event exchange.
    aud = rt.audio_stream
    main = torch.cuda.current_stream() if aud is not None else None

    def on_audio():
        return torch.cuda.stream(aud) if aud is not None else contextlib.nullcontext()

    def record(stream):
        if aud is None:
            return None
        ev = torch.cuda.Event()
        ev.record(stream)
        return ev

    def wait(stream, ev):
        if ev is not None:
            stream.wait_event(ev)

    if aud is not None:
        # everything the audio stream will touch is built on the main stream first: packed weights re-point (and
        # free) the original parameters, which belong to the main stream's allocator pool
        prepack(audio_dit)
        prepack(bridge)
        aud.wait_stream(main)
    ev_a = None  # audio_x of this layer is complete on the audio stream
    for i in range(min_layers):
        ev_v = record(main)  # x_loc of this layer is complete on the main stream
        wait(main, ev_a)
        x_pre, a_pre = x_loc, audio_x
        if aud is not None:  # tensors read on a stream other than the one that allocated them
            a_pre.record_stream(main)
            x_pre.record_stream(aud)
        a2v = bridge.should_interact(i, "a2v")
        if a2v:
            # a2v: local video queries x replicated audio keys -- no communication
            with _seg("bridge_a2v"):
                x_loc = bridge.audio_to_video_conditioners[str(i)].forward_residual(
                    x_pre, a_pre, v_cs_loc, a_cs, scale_for(a2v_condition_scale))
        with on_audio():
            wait(aud, ev_v)
            side = "audio" if aud is not None else "main"
            if a2v and bridge.should_interact(i, "v2a"):
                with _seg("bridge_v2a(+all_gathers)", side):
                    audio_x = _v2a_cp(bridge.video_to_audio_conditioners[str(i)], a_pre, x_pre, a_cs, v_cs_loc,
                                      scale_for(v2a_condition_scale), rt)
            with _seg("audio_block", side):
                audio_x = audio_dit.blocks[i](audio_x, audio_context, audio_t_mod, a_tab)  # replicated
            ev_a = record(aud)
        x_loc = _video_block_cp(visual_dit.blocks[i], x_loc, visual_context, visual_t_mod, v_tab_loc, rt, rows)
    wait(main, ev_a)
    if aud is not None:
        audio_x.record_stream(main)
    for i in range(min_layers, visual_layers):
        x_loc = _video_block_cp(visual_dit.blocks[i], x_loc, visual_context, visual_t_mod, v_tab_loc, rt, rows)
    if not _gather:
        return x_loc, audio_x, rows, rt.group
    visual_full = cpmod.all_gather_cat(x_loc, rows, rt.group, dim=1)
    return visual_full, audio_x


class GraphedForward:
    """CUDA-graph replay of the whole forward for one set of argument shapes.

    One forward is ~1600 kernel launches; the 403-token audio tower and the bridge's small kernels are launch-bound
    (each runs for a few microseconds), which shows at cp = 8 where a video layer is down to ~8 ms.  Capturing the
    launches once (TMA descriptors, NCCL all-to-alls on the side stream and all) and replaying the graph removes the
    host from the loop.  Inputs are copied into static buffers before the replay, outputs are cloned after it, so
    the caller sees ordinary tensors; RoPE tables are rebuilt inside the graph from the (static) complex inputs."""

    def __init__(self, owner, visual_dit, tensors: dict, statics: dict, cp_mesh):
        self.static_in = {k: v.clone() for k, v in tensors.items()}
        self.statics = statics
        self.cp_mesh = cp_mesh
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        rope.set_cache(False)  # table conversions must be recorded as graph nodes, not served from the memo
        try:
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up: packs weights, configures kernels, builds the NCCL channels
                    _forward_eager(owner, visual_dit, **self.static_in, **statics, cp_mesh=cp_mesh)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_out = _forward_eager(owner, visual_dit, **self.static_in, **statics, cp_mesh=cp_mesh)
        finally:
            rope.set_cache(True)

    def __call__(self, tensors: dict):
        for k, v in tensors.items():
            self.static_in[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return tuple(t.clone() for t in self.static_out)


_TENSOR_ARGS = ("visual_x", "audio_x", "visual_context", "audio_context", "visual_t_mod", "audio_t_mod", "visual_freqs",
                "audio_freqs")


@torch.no_grad()
def forward_dual_tower_dit(self, visual_dit, visual_x: torch.Tensor, audio_x: torch.Tensor,
                           visual_context: torch.Tensor, audio_context: torch.Tensor, visual_t_mod: torch.Tensor,
                           audio_t_mod: Optional[torch.Tensor], visual_freqs: torch.Tensor, audio_freqs: torch.Tensor,
                           grid_size: Tuple[int, int, int], video_fps: float, condition_scale: Optional[float] = 1.0,
                           a2v_condition_scale: Optional[float] = None, v2a_condition_scale: Optional[float] = None,
                           cp_mesh=None):
    """Drop-in for ``MOVA.forward_dual_tower_dit`` (pipeline_mova.py:612-711), same arguments and return value.
    Runs eagerly, or -- when ``install(pipe, cuda_graph=True)`` / ``pipe.mova_b200_cuda_graph = True`` -- as a CUDA
    graph captured on first use per argument-shape signature."""
    statics = dict(grid_size=tuple(int(g) for g in grid_size), video_fps=float(video_fps),
                   condition_scale=condition_scale, a2v_condition_scale=a2v_condition_scale,
                   v2a_condition_scale=v2a_condition_scale)
    tensors = dict(visual_x=visual_x, audio_x=audio_x, visual_context=visual_context, audio_context=audio_context,
                   visual_t_mod=visual_t_mod, audio_t_mod=audio_t_mod, visual_freqs=visual_freqs, audio_freqs=audio_freqs)
    if not getattr(self, "mova_b200_cuda_graph", False):
        return _forward_eager(self, visual_dit, **tensors, **statics, cp_mesh=cp_mesh)
    if any(hasattr(m, "_hf_hook") for m in (visual_dit, self.audio_dit, self.dual_tower_bridge)):
        raise RuntimeError("dualforce_b200: cuda_graph=True bakes weight pointers into the graph; it cannot be combined "
                           "with accelerate CPU-offload hooks (weights move between forwards)")
    # a captured graph holds raw pointers: scales must be plain floats BEFORE capture (a 0-dim Parameter would need
    # .item(), a host sync that aborts capture), and the key must change when the weights move
    bridge = self.dual_tower_bridge
    for name in ("condition_scale", "a2v_condition_scale", "v2a_condition_scale"):
        if isinstance(statics[name], torch.Tensor):
            statics[name] = bridge._scale(statics[name])
    if statics["condition_scale"] is None or isinstance(bridge.condition_scale, torch.Tensor):
        statics["condition_scale"] = bridge._scale(statics["condition_scale"])
    blocks = visual_dit.blocks
    a2v = list(bridge.audio_to_video_conditioners.values())
    weights_sig = param_sig(blocks[0].self_attn.q.weight, blocks[len(blocks) - 1].ffn[2].weight,
                            self.audio_dit.blocks[0].self_attn.q.weight, a2v[0].inner.q.weight if a2v else None)
    cache = self.__dict__.setdefault("_mova_b200_graphs", {})
    key = (id(visual_dit), id(cp_mesh), tuple(sorted((k, v) for k, v in statics.items())), weights_sig,
           tuple((k, tuple(v.shape), v.dtype, str(v.device)) for k, v in tensors.items()))
    runner = cache.get(key)
    if runner is None:
        runner = cache[key] = GraphedForward(self, visual_dit, tensors, statics, cp_mesh)
    return runner(tensors)


# ----------------------------------------------------------------------------------------------------------------
# installer
# ----------------------------------------------------------------------------------------------------------------
def _swap_blocks(model) -> int:
    n = 0
    for i, blk in enumerate(model.blocks):
        if not isinstance(blk, DiTBlock):
            model.blocks[i] = DiTBlock.from_reference(blk)
            n += 1
    return n


def install(pipe, cuda_graph: bool = False, bridge_rope: str = "reference_bf16") -> int:
    """Swap the B200 modules into a reference ``MOVA`` pipeline (or any object with ``video_dit``,
    ``video_dit_2``, ``audio_dit``, ``dual_tower_bridge``) in place, sharing its parameters, and bind
    ``pipe.forward_dual_tower_dit`` to the B200 path.  Returns the number of modules replaced -- the same idiom as
    ``MOVA.replace_attention`` (pipeline_mova.py:124-148), which becomes unnecessary (and must not be called after
    this): context parallelism is handled inside ``forward_dual_tower_dit`` from ``cp_mesh``.

    ``bridge_rope``: precision of the bridge's aligned cross-RoPE tables.  ``"reference_bf16"`` (default) reproduces
    what the reference computes when it runs as shipped -- bf16-rounded ``inv_freq`` and bf16 cos / sin tables
    (interactionv2.py:21-23, 36), the only tables a MOVA checkpoint has been trained and sampled with; ``"fp32"``
    keeps exact frequencies (what the reference computes in fp32; the CPU oracle's default).  See
    ``modules.RotaryEmbedding.forward`` and INTEGRATION.md.

    Raises if the extension library is missing or the device is not sm_100 -- there is no fallback."""
    from . import _lib

    _lib.load()
    if torch.cuda.is_available():
        _lib.require_device(torch.cuda.current_device())
    else:
        raise _lib.MovaB200Error("dualforce_b200.install: no CUDA device (the B200 path has no CPU fallback)")
    count = swap_modules(pipe)
    pipe.mova_b200_cuda_graph = bool(cuda_graph)
    if getattr(pipe, "dual_tower_bridge", None) is not None:
        pipe.dual_tower_bridge.bridge_rope = bridge_rope
    return count


def swap_modules(pipe) -> int:
    """The host-side half of :func:`install` (no device checks): replace DiTBlocks, bridge and heads by their B200
    twins sharing the reference Parameters, bind ``forward_dual_tower_dit`` and ``inference_single_step``."""
    from . import step

    count = 0
    for name in ("video_dit", "video_dit_2", "audio_dit"):
        model = getattr(pipe, name, None)
        if model is not None:
            count += _swap_blocks(model)
    bridge = getattr(pipe, "dual_tower_bridge", None)
    if bridge is not None and not isinstance(bridge, DualTowerConditionalBridge):
        pipe.dual_tower_bridge = DualTowerConditionalBridge.from_reference(bridge)
        count += len(pipe.dual_tower_bridge.audio_to_video_conditioners) + len(
            pipe.dual_tower_bridge.video_to_audio_conditioners)
    pipe.forward_dual_tower_dit = types.MethodType(forward_dual_tower_dit, pipe)
    if not hasattr(pipe, "mova_b200_cuda_graph"):
        pipe.mova_b200_cuda_graph = False
    # the step around the path (pipeline_mova.py:500-609): heads share the reference parameters
    for name in ("video_dit", "video_dit_2", "audio_dit"):
        model = getattr(pipe, name, None)
        if model is not None and hasattr(model, "head") and not isinstance(model.head, step.Head):
            model.head = step.Head.from_reference(model.head)
            count += 1
    step.bind(pipe)
    # scripts/inference_single.py:102-117 calls pipe.replace_attention() when cp_size > 1; on an installed pipeline that
    # must not put the reference's yunchang-backed USPAttention back into the modules (context parallelism is handled
    # from cp_mesh here), so the method keeps its contract -- it returns the number of attention sites -- and leaves
    # the B200 processors where they are.
    pipe.replace_attention = types.MethodType(_replace_attention_noop, pipe)
    return count


def _replace_attention_noop(self, attn_type=None) -> int:
    """``MOVA.replace_attention`` (pipeline_mova.py:124-148) on an installed pipeline: counts the same sites the
    reference would replace and replaces nothing."""
    n = 0
    for name in ("video_dit", "video_dit_2", "audio_dit"):
        model = getattr(self, name, None)
        if model is not None:
            n += len(model.blocks)
    bridge = getattr(self, "dual_tower_bridge", None)
    if bridge is not None:
        n += len(bridge.audio_to_video_conditioners) + len(bridge.video_to_audio_conditioners)
    return n
