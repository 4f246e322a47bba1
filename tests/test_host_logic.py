"""CPU tests of the host side: RoPE table plumbing, bridge tables and controller, state-dict compatibility with the
reference, parameter packing / sharing.  No kernel is launched."""
import json
import os

import numpy as np
import pytest
import torch

import mova_oracle as O
import ref_loader

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def meta():
    with open(os.path.join(GOLDEN, "tiny_dual_tower.json")) as f:
        return json.load(f)


def test_rope_tables_match_oracle():
    from dualforce_b200 import rope

    grid = (3, 4, 5)
    v = rope.video_freqs(rope.precompute_freqs_cis_3d(128), grid, "cpu")
    assert torch.equal(v, O.video_freqs(128, grid))
    a = rope.audio_freqs(rope.precompute_freqs_cis_1d(128), 21, "cpu")
    assert torch.equal(a, O.audio_freqs(128, 21))
    cos, sin = rope.as_tables(v)
    assert cos.dtype == torch.float32 and cos.shape == (60, 64) and cos.is_contiguous()
    assert torch.equal(cos, v.real.float().reshape(60, 64)) and torch.equal(sin, v.imag.float().reshape(60, 64))
    assert rope.as_tables(v)[0] is cos  # memoised on the identity of the source tensor
    v.mul_(1.0)  # an in-place write bumps the version: the memo must not serve the old table
    assert rope.as_tables(v)[0] is not cos
    rope.set_cache(False)
    try:
        assert rope.as_tables(v)[0] is not rope.as_tables(v)[0]
    finally:
        rope.set_cache(True)
    rope.clear_cache()
    with pytest.raises(TypeError):
        rope.tables_from_complex(torch.zeros(4, 64))


def test_bridge_tables_and_controller_match_reference_golden(meta):
    import dualforce_b200 as B

    cfg = meta["cfg"]
    gold = np.load(os.path.join(GOLDEN, "tiny_dual_tower.npz"))
    bridge = B.DualTowerConditionalBridge(
        visual_layers=cfg["visual_layers"], audio_layers=cfg["audio_layers"], visual_hidden_dim=cfg["visual_dim"],
        audio_hidden_dim=cfg["audio_dim"], audio_fps=cfg["audio_fps"], head_dim=cfg["head_dim"],
        interaction_strategy=cfg["interaction_strategy"], apply_cross_rope=True)
    (cv, sv), (ca, sa) = bridge.build_aligned_freqs(cfg["video_fps"], tuple(cfg["grid_size"]), cfg["audio_len"],
                                                    device=torch.device("cpu"), dtype=torch.float32)
    for got, key in ((cv, "cos_v"), (sv, "sin_v"), (ca, "cos_a"), (sa, "sin_a")):
        assert (got - torch.from_numpy(gold[key])).abs().max() < 1e-6
    # memoised on the arguments, and immune to the bf16 cast of the module (deviation 2)
    again = bridge.build_aligned_freqs(cfg["video_fps"], tuple(cfg["grid_size"]), cfg["audio_len"],
                                       device=torch.device("cpu"), dtype=torch.float32)
    assert again[0][0] is cv
    bridge.to(torch.bfloat16)
    bridge._freq_cache.clear()
    (cv2, _), _ = bridge.build_aligned_freqs(cfg["video_fps"], tuple(cfg["grid_size"]), cfg["audio_len"],
                                             device=torch.device("cpu"), dtype=torch.float32)
    assert torch.equal(cv2, cv)
    # controller strategies
    for strategy in ("shallow_focus", "distributed", "progressive", "custom", "full"):
        for nv, na in ((40, 30), (30, 30), (3, 2), (12, 20)):
            ctl = B.CrossModalInteractionController(nv, na)
            layers = [i for i, _ in ctl.get_interaction_layers(strategy)["a2v"]]
            assert layers == O.interaction_layers(strategy, nv, na)
    with pytest.raises(ValueError):
        B.CrossModalInteractionController(3, 3).get_interaction_layers("nope")
    assert bridge.should_interact(0, "a2v") and bridge.should_interact(1, "v2a") and not bridge.should_interact(2, "a2v")


def test_state_dict_keys_equal_the_reference(meta):
    import dualforce_b200 as B

    cfg = meta["cfg"]
    blk = B.DiTBlock(False, cfg["visual_dim"], cfg["visual_heads"], cfg["visual_ffn"], cfg["eps"])
    assert sorted(blk.state_dict().keys()) == meta["reference_state_dict_keys"]["dit_block"]
    bridge = B.DualTowerConditionalBridge(
        visual_layers=cfg["visual_layers"], audio_layers=cfg["audio_layers"], visual_hidden_dim=cfg["visual_dim"],
        audio_hidden_dim=cfg["audio_dim"], audio_fps=cfg["audio_fps"], head_dim=cfg["head_dim"],
        interaction_strategy=cfg["interaction_strategy"], apply_cross_rope=True)
    assert sorted(bridge.state_dict().keys()) == meta["reference_state_dict_keys"]["bridge"]
    # the oracle's synthetic weights load strictly (same keys, same shapes)
    Pv, Pa, Pb, _ = O.make_case(dict(cfg, grid_size=tuple(cfg["grid_size"])), 1)
    blk.load_state_dict({k[len("blocks.0."):]: v for k, v in Pv.items() if k.startswith("blocks.0.")}, strict=True)
    bridge.load_state_dict(Pb, strict=True)


def test_unsupported_reference_options_fail_loudly():
    import dualforce_b200 as B

    with pytest.raises(NotImplementedError):
        B.DiTBlock(True, 256, 2, 512)
    with pytest.raises(NotImplementedError):
        B.ConditionalCrossAttentionBlock(256, 128, 2, pooled_adaln=True)
    with pytest.raises(NotImplementedError):
        B.DualTowerConditionalBridge(head_dim=64)
    with pytest.raises(RuntimeError):
        B.GateModule()(None, None, None)


def test_packing_keeps_state_dict_and_shares_memory():
    from dualforce_b200.modules import SelfAttention, _Packed

    sa = SelfAttention(256, 2)
    before = {k: v.clone() for k, v in sa.state_dict().items()}
    packed = _Packed()
    w, b = packed.get([sa.q, sa.k, sa.v])
    assert w.shape == (768, 256) and b.shape == (768,)
    after = sa.state_dict()
    assert set(after) == set(before) and all(torch.equal(after[k], before[k]) for k in before)
    assert sa.k.weight.data_ptr() == w[256:512].data_ptr()  # parameters are views of the packed buffer
    with torch.no_grad():
        sa.v.bias.add_(1.0)  # in-place updates (load_state_dict) stay visible to the GEMM
    assert torch.equal(b[512:], sa.v.bias.data)
    assert packed.get([sa.q, sa.k, sa.v])[0] is w  # no repack while the parameters have not moved
    sa.to(torch.float64)
    assert packed.get([sa.q, sa.k, sa.v])[0] is not w  # .to() re-materialised the parameters: repack


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_from_reference_shares_parameters_and_counts_like_replace_attention():
    import types

    import dualforce_b200 as B
    from dualforce_b200.pipeline import _swap_blocks

    R = ref_loader.load()
    ref_blocks = torch.nn.ModuleList([R.wan_video_dit.DiTBlock(False, 256, 2, 512, 1e-6) for _ in range(3)])
    model = types.SimpleNamespace(blocks=ref_blocks)
    ref0 = ref_blocks[0]
    assert _swap_blocks(model) == 3 and _swap_blocks(model) == 0
    new0 = model.blocks[0]
    assert isinstance(new0, B.DiTBlock)
    assert new0.self_attn.q.weight is ref0.self_attn.q.weight and new0.modulation is ref0.modulation
    assert sorted(new0.state_dict().keys()) == sorted(ref0.state_dict().keys())
    ref_bridge = R.interactionv2.DualTowerConditionalBridge(visual_layers=3, audio_layers=2, visual_hidden_dim=256,
                                                            audio_hidden_dim=128, audio_fps=50.0, head_dim=128,
                                                            interaction_strategy="full", apply_cross_rope=True)
    br = B.DualTowerConditionalBridge.from_reference(ref_bridge)
    assert sorted(br.state_dict().keys()) == sorted(ref_bridge.state_dict().keys())
    assert br.audio_to_video_conditioners["1"].inner.k.weight is ref_bridge.audio_to_video_conditioners["1"].inner.k.weight
    assert br.apply_cross_rope and br.should_interact(1, "v2a")


def test_sp_helper_twins_match_reference_golden(meta):
    """dualforce_b200.cp._sp_* (forward-only twins of mova/distributed/functional.py) against the reference's own
    outputs in the golden file: pure index work, bit-exact."""
    import numpy as np

    import mova_oracle as O
    from dualforce_b200 import cp

    gold = np.load(os.path.join(GOLDEN, "tiny_dual_tower.npz"))
    cfg = dict(meta["cfg"], grid_size=tuple(meta["cfg"]["grid_size"]))
    _, _, _, inp = O.make_case(cfg, meta["seed"])
    x = inp["audio_x"]  # [1, 21, 128]: 4 ranks -> chunks of 6, last one short, 3 pad rows
    chunks = []
    for r in range(4):
        c, chunk_len, pad_len, total = cp._sp_split_tensor(x, sp_size=4, sp_rank=r)
        assert torch.equal(c, torch.from_numpy(gold[f"sp_split_r{r}"]))
        assert (chunk_len, pad_len, total) == (6, 3, 21)
        chunks.append(c)
    full = torch.cat(chunks, dim=1)[:, :-3]
    assert torch.equal(full, x)
    # surplus rank (more ranks than chunks) gets zeros; dim-0 variant on a table
    c, chunk_len, pad_len, total = cp._sp_split_tensor(x[:, :3], sp_size=4, sp_rank=3)
    assert chunk_len == 1 and pad_len == 1 and torch.equal(c, torch.zeros(1, 1, 128))
    tab = torch.arange(21 * 4, dtype=torch.float32).reshape(21, 1, 4)
    c0, cl, pl, tot = cp._sp_split_tensor_dim_0(tab, sp_size=4, sp_rank=3)
    assert (cl, pl, tot) == (6, 3, 21) and torch.equal(c0[:3], tab[18:]) and torch.equal(c0[3:], torch.zeros(3, 1, 4))
    sel = cp._sp_select_rank(x, sp_size=4, sp_rank=3, chunk_len=6, pad_len=3)
    assert torch.equal(sel, chunks[3])


def test_activation_hook_installs_on_first_call(monkeypatch):
    """dualforce_b200.launch.activate: the wrapped __call__ installs once, then defers to the original."""
    import dualforce_b200
    from dualforce_b200 import launch

    calls = []

    class FakeMOVA:
        def __call__(self, prompt, steps=1):
            calls.append(("call", prompt, steps))
            return "video"

    monkeypatch.setattr(dualforce_b200, "install", lambda pipe, cuda_graph=False: calls.append(("install", cuda_graph)) or 0)
    cls = launch.activate(FakeMOVA)
    assert launch.activate(FakeMOVA) is cls  # idempotent
    pipe = FakeMOVA()
    assert pipe("a", steps=2) == "video" and pipe("b") == "video"
    assert calls == [("install", False), ("call", "a", 2), ("call", "b", 1)]
    other = FakeMOVA()
    monkeypatch.setenv("MOVA_B200_CUDA_GRAPH", "1")
    other("c")
    assert calls[-2:] == [("install", True), ("call", "c", 1)]


def test_peer_push_is_dealt_by_destination():
    """dualforce_b200.peer: with k push streams every destination's copies and its flag stay on ONE stream, in order,
    and nothing is lost or duplicated; flag indices of the exchange are disjoint per (direction, group, source)."""
    from dualforce_b200 import peer

    order = [3, 4, 5, 6, 7, 0, 1]  # rank 2 of 8: remote destinations in push order
    copies = [(d, 100 * d, f"chunk{d}") for d in order]
    flags = [(d, 17) for d in order]
    for k in (1, 2, 3, 4, 7, 8):
        parts = peer.deal_by_destination(copies, flags, k)
        assert len(parts) == min(k, len(order))
        seen = []
        for cj, fj in parts:
            assert [c[0] for c in cj] == [f[0] for f in fj]  # a destination's data and flag travel together
            seen += [c[0] for c in cj]
        assert sorted(seen) == sorted(order)
        assert [c[0] for c in parts[0][0]] == order[0::k]
    # local pushes (one destination, several groups) stay one push
    local = peer.deal_by_destination([(2, g, None) for g in range(5)], [(2, g) for g in range(5)], 4)
    assert len(local) == 1 and len(local[0][0]) == 5 and len(local[0][1]) == 5
    idx = {peer.PeerExchange.flag_index(d, g, s, 8) for d in (0, 1) for g in range(peer.MAX_GROUPS) for s in range(8)}
    assert len(idx) == 2 * peer.MAX_GROUPS * 8 and max(idx) < peer.EPOCH_SRC_OFFSET // 8
