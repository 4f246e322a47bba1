"""Context-parallel host logic on CPU: world_size 2 over gloo.  The kernels are emulated with plain torch / the
oracle; what is under test is the layout and collective plumbing of dualforce_b200/cp.py (segment order, weight
permutations, ragged all-to-all splits, LSE merge), i.e. everything the N>1 path adds around the kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mova_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, L, H, groups, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dualforce_b200 import cp

        torch.manual_seed(0)  # same tensors on every rank
        D = 128
        d = H * D
        x = torch.randn(1, L, d)
        P = O.make_block_weights(torch.Generator().manual_seed(3), "b", d, 2 * d)
        freqs = O.video_freqs(D, (L, 1, 1))
        # reference: unsharded self-attention (oracle)
        ref = O.self_attention(P, "b.self_attn", x, freqs, H, 1e-6)

        chunks = cp.seq_chunks(L, world)
        rows = [b - a for a, b in chunks]
        s0, s1 = chunks[rank]
        plan = cp.UlyssesPlan(H, D, world, groups)
        Wqkv = torch.cat([P["b.self_attn.q.weight"], P["b.self_attn.k.weight"], P["b.self_attn.v.weight"]], 0)
        bqkv = torch.cat([P["b.self_attn.q.bias"], P["b.self_attn.k.bias"], P["b.self_attn.v.bias"]], 0)
        ridx, cidx = plan.qkv_row_index(), plan.channel_index()
        h = x[0, s0:s1]
        Lc = s1 - s0
        # "QKV GEMM with out_segments": C[m, n] with permuted rows of W, stored [nseg, Lc, 3w]
        qkv = (h @ Wqkv[ridx].t() + bqkv[ridx]).reshape(Lc, plan.nseg, 3 * plan.w).permute(1, 0, 2).contiguous()
        # "segmented RMSNorm + RoPE": the logical row is the concatenation of the q parts of all segments
        for part, nw in ((0, "b.self_attn.norm_q.weight"), (1, "b.self_attn.norm_k.weight")):
            row = qkv[:, :, part * plan.w:(part + 1) * plan.w].permute(1, 0, 2).reshape(1, Lc, d)  # segment order
            row = O.rms_norm(row, P[nw][cidx], 1e-6)
            row = O.rope_interleaved(row, freqs[s0:s1], D)  # per-head rotation: order of heads is irrelevant
            qkv[:, :, part * plan.w:(part + 1) * plan.w] = row.reshape(Lc, plan.nseg, plan.w).permute(1, 0, 2)
        send = qkv.view(groups, world, Lc, 3 * plan.w)
        back = torch.empty(groups, world, Lc, plan.w)
        for g in range(groups):
            recv = cp.scatter_heads(send[g], rows, rank, None)  # [L, 3w]
            assert recv.shape == (L, 3 * plan.w)
            w = plan.w
            o = O.attention(recv[None, :, :w], recv[None, :, w:2 * w], recv[None, :, 2 * w:], plan.Hg)[0]
            cp.gather_heads(o.contiguous(), rows, rank, None, out=back[g])
        # "o-proj with segmented A": A[m, seg*w + c] = back[seg, m, c], weight columns permuted
        A = back.view(plan.nseg, Lc, plan.w).permute(1, 0, 2).reshape(Lc, d)
        out = A @ P["b.self_attn.o.weight"][:, cidx].t() + P["b.self_attn.o.bias"]
        err = (out - ref[0, s0:s1]).abs().max().item()
        # final gather of ragged chunks
        full = cp.all_gather_cat(out[None], rows, None, dim=1)
        err_full = (full - ref).abs().max().item()
        # v2a-style partial attention over this rank's keys + exact merge
        q = torch.randn(1, 5, d)
        o_p, lse_p = O.attention(q, x[:, s0:s1], x[:, s0:s1] * 0.5, H, return_lse=True)
        o_all = cp.all_gather_stack(o_p[0], world, None)
        lse_all = cp.all_gather_stack(lse_p[0], world, None)
        merged = O.merge_partial_attention([(o_all[r][None], lse_all[r][None]) for r in range(world)])
        err_merge = (merged - O.attention(q, x, x * 0.5, H)).abs().max().item()
        results[rank] = (err, err_full, err_merge)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("L,H,groups", [(12, 4, 1), (11, 4, 2), (7, 2, 1)])
def test_ulysses_layouts_world2(L, H, groups):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), L, H, groups, results), nprocs=2, join=True)
    assert len(results) == 2
    for rank, (err, err_full, err_merge) in results.items():
        assert err < 2e-4, f"rank {rank}: sharded self-attention differs from unsharded by {err}"
        assert err_full < 2e-4
        assert err_merge < 1e-5


def test_plan_indices_are_permutations():
    from dualforce_b200 import cp

    for H, c, g in ((40, 8, 1), (40, 4, 2), (40, 2, 2), (12, 4, 1)):
        plan = cp.UlyssesPlan(H, 128, c, g)
        assert sorted(plan.channel_index().tolist()) == list(range(H * 128))
        assert sorted(plan.qkv_row_index().tolist()) == list(range(3 * H * 128))
        assert plan.nseg * plan.w == H * 128
    assert cp.UlyssesPlan.pick_groups(5) == 1 and cp.UlyssesPlan.pick_groups(10) == 2 and cp.UlyssesPlan.pick_groups(20) == 2
    assert cp.seq_chunks(43120, 8) == [(i * 5390, (i + 1) * 5390) for i in range(8)]
    assert cp.seq_chunks(11, 2) == [(0, 6), (6, 11)]
    with pytest.raises(ValueError):
        cp.seq_chunks(3, 8)


def _usp_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dualforce_b200 import cp

        torch.manual_seed(1)
        H, D, S, Skv, B = 4, 16, 10, 6, 2
        q, k, v = torch.randn(B, S, H * D), torch.randn(B, Skv, H * D), torch.randn(B, Skv, H * D)
        ref = O.attention(q, k, v, H)
        sl, kl = S // world, Skv // world
        out = cp.ulysses_attention(q[:, rank * sl:(rank + 1) * sl].contiguous(), k[:, rank * kl:(rank + 1) * kl].contiguous(),
                                   v[:, rank * kl:(rank + 1) * kl].contiguous(), H,
                                   lambda a, b, c, h: O.attention(a, b, c, h))
        results[rank] = (out - ref[:, rank * sl:(rank + 1) * sl]).abs().max().item()
    finally:
        dist.destroy_process_group()


def test_usp_attention_processor_world2():
    """The USPAttention drop-in (reference contract: sequence shards in, sequence shards out) with the kernel replaced
    by the oracle: head <-> sequence all-to-all and its inverse."""
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_usp_worker, args=(2, _free_port(), results), nprocs=2, join=True)
    assert len(results) == 2 and all(e < 1e-5 for e in results.values()), dict(results)


def _peer_exchange_worker(rank, world, port, results):
    """dualforce_b200.peer.PeerExchange on its own (shared-memory windows): ragged rows, several head groups, a second
    round with larger shapes that forces the windows to grow (collective re-allocation), flags by epoch."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dualforce_b200 import peer
        from shm_peer import ShmWindow

        win = ShmWindow(rank, world, f"px{port}", jitter_ms=2.0 * (1 + rank))
        px = peer.PeerExchange(win, rank, world)
        ok = True
        for rnd, (G, rows, w) in enumerate([(2, [5, 3, 4][:world] if world == 3 else [5, 3][:world], 128),
                                            (3, [70, 64, 61][:world] if world == 3 else [70, 64][:world], 256)]):
            L, C, Lc = sum(rows), 3 * w, rows[rank]
            off = [sum(rows[:r]) for r in range(world)]
            recv, back = px.begin(G, L, C, rows, w)
            assert recv.shape == (G, L, C) and back.shape == (G, world, Lc, w) and px.epoch == rnd + 1
            # value of element = f(group, source rank, destination rank, row, column): checkable on the other side
            def chunk(g, src, dst, nrows, ncols):
                base = 1000 * g + 100 * src + 10 * dst
                return (base + torch.arange(nrows * ncols).reshape(nrows, ncols) % 7).to(torch.bfloat16)
            send = torch.stack([torch.stack([chunk(g, rank, d, Lc, C) for d in range(world)]) for g in range(G)])
            px.push_in_local(send)
            for g in range(G):
                px.push_in(g, send[g])
            px.wait_in(0, G - 1)
            for g in range(G):
                for s in range(world):
                    ok &= bool(torch.equal(recv[g, off[s]:off[s] + rows[s]], chunk(g, s, rank, rows[s], C)))
            # the way back: rows of rank d of my [L, w] output go to rank d's back[g, me]
            outs = [torch.cat([chunk(g, rank, d, rows[d], w) + 1 for d in range(world)]) for g in range(G)]
            for g in range(G):
                px.push_out(g, outs[g])
            px.push_out_local([(g, outs[g]) for g in range(G)])
            px.wait_out()
            for g in range(G):
                for s in range(world):
                    ok &= bool(torch.equal(back[g, s], chunk(g, s, rank, Lc, w) + 1))
            dist.barrier()  # a forward always ends with a collective before shapes may change (peer.py invariants)
        results[rank] = dict(ok=ok, generation=win.generation, epoch=px.epoch, capacity=win.capacity)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_exchange_rounds_ragged_and_growing(world):
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_peer_exchange_worker, args=(world, port, results), nprocs=world, join=True)
    assert len(results) == world
    for rank in range(world):
        r = results[rank]
        assert r["ok"], f"rank {rank}: exchanged bytes differ"
        assert r["generation"] == 2 and r["epoch"] == 2, r  # the second round did not fit the first windows
