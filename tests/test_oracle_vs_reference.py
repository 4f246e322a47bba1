"""The oracle against the live reference modules (only where /root/reference is mounted, i.e. the authoring
container): a second seed and a second geometry, beyond the committed golden vectors."""
import pytest
import torch

import mova_oracle as O
import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("seed,grid,audio_len", [(7, (2, 3, 4), 10), (99, (4, 2, 3), 33)])
def test_full_path_matches_reference(seed, grid, audio_len):
    import make_golden

    cfg = dict(O.TINY_CFG, grid_size=grid, audio_len=audio_len)
    arrays, _, _ = make_golden.run_reference(cfg, seed)
    Pv, Pa, Pb, inp = O.make_case(cfg, seed)
    fv, fa = O.forward_dual_tower_dit(Pv, Pa, Pb, cfg, inp["visual_x"], inp["audio_x"], inp["visual_context"],
                                      inp["audio_context"], inp["visual_t_mod"], inp["audio_t_mod"], inp["visual_freqs"],
                                      inp["audio_freqs"], cfg["grid_size"], cfg["video_fps"])
    for got, key in ((fv, "final_visual"), (fa, "final_audio")):
        ref = torch.from_numpy(arrays[key])
        assert (got - ref).abs().max() <= 5e-5 * max(ref.abs().max().item(), 1.0)


def test_rope_tables_match_reference():
    R = ref_loader.load()
    f3 = R.wan_video_dit.precompute_freqs_cis_3d(128)
    f, h, w = 3, 4, 5
    ref = torch.cat([f3[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1), f3[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                     f3[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(f * h * w, 1, -1)
    assert torch.equal(O.video_freqs(128, (f, h, w)), ref)
    fa = R.wan_audio_dit.precompute_freqs_cis_1d(128)
    L = 21
    ref_a = torch.cat([fa[0][:L].view(L, -1), fa[1][:L].view(L, -1), fa[2][:L].view(L, -1)], dim=-1).reshape(L, 1, -1)
    assert torch.equal(O.audio_freqs(128, L), ref_a)


def test_first_frame_bias_tables_match_reference():
    R = ref_loader.load()
    bridge = R.interactionv2.DualTowerConditionalBridge(visual_layers=2, audio_layers=2, visual_hidden_dim=256,
                                                        audio_hidden_dim=128, audio_fps=50.0, head_dim=128,
                                                        interaction_strategy="full", apply_cross_rope=True,
                                                        apply_first_frame_bias_in_rope=True)
    ref_v, ref_a = bridge.build_aligned_freqs(video_fps=24.0, grid_size=(5, 2, 3), audio_steps=17,
                                              device=torch.device("cpu"), dtype=torch.float32)
    got_v, got_a = O.build_aligned_freqs(24.0, (5, 2, 3), 17, 50.0, 128, apply_first_frame_bias=True)
    for g, r in zip(got_v + got_a, ref_v + ref_a):
        assert (g - r).abs().max() < 1e-6
    # and the B200 module's own builder (pure torch, runs on CPU)
    import dualforce_b200 as B

    mine = B.DualTowerConditionalBridge(visual_layers=2, audio_layers=2, visual_hidden_dim=256, audio_hidden_dim=128,
                                        audio_fps=50.0, head_dim=128, interaction_strategy="full", apply_cross_rope=True,
                                        apply_first_frame_bias_in_rope=True)
    m_v, m_a = mine.build_aligned_freqs(24.0, (5, 2, 3), 17, device=torch.device("cpu"), dtype=torch.float32)
    for g, r in zip(m_v + m_a, ref_v + ref_a):
        assert (g - r).abs().max() < 1e-6


@pytest.mark.parametrize("strategy,nv,na", [("shallow_focus", 7, 6), ("distributed", 5, 4), ("custom", 4, 3)])
def test_sparse_interaction_strategies_and_scales_match_reference(strategy, nv, na):
    """Bridge only on some layers, different a2v / v2a strengths, more video than audio layers."""
    import types

    cfg = dict(O.TINY_CFG, visual_layers=nv, audio_layers=na, interaction_strategy=strategy, grid_size=(2, 2, 3),
               audio_len=9, visual_dim=128, visual_heads=1, visual_ffn=128, audio_dim=128, audio_ffn=128, text_len=4)
    Pv, Pa, Pb, inp = O.make_case(cfg, 5)
    R = ref_loader.load()
    DiTBlock = R.wan_video_dit.DiTBlock
    vis, aud = torch.nn.Module(), torch.nn.Module()
    vis.blocks = torch.nn.ModuleList([DiTBlock(False, 128, 1, 128, 1e-6) for _ in range(nv)])
    aud.blocks = torch.nn.ModuleList([DiTBlock(False, 128, 1, 128, 1e-6) for _ in range(na)])
    bridge = R.interactionv2.DualTowerConditionalBridge(visual_layers=nv, audio_layers=na, visual_hidden_dim=128,
                                                        audio_hidden_dim=128, audio_fps=50.0, head_dim=128,
                                                        interaction_strategy=strategy, apply_cross_rope=True)
    vis.load_state_dict(Pv, strict=True)
    aud.load_state_dict(Pa, strict=True)
    bridge.load_state_dict(Pb, strict=True)
    pipe = types.SimpleNamespace(audio_dit=aud, dual_tower_bridge=bridge)
    with torch.no_grad():
        rv, ra = R.forward_dual_tower_dit(
            pipe, visual_dit=vis, visual_x=inp["visual_x"], audio_x=inp["audio_x"],
            visual_context=inp["visual_context"], audio_context=inp["audio_context"], visual_t_mod=inp["visual_t_mod"],
            audio_t_mod=inp["audio_t_mod"], visual_freqs=inp["visual_freqs"], audio_freqs=inp["audio_freqs"],
            grid_size=cfg["grid_size"], video_fps=24.0, condition_scale=0.7)
    fv, fa = O.forward_dual_tower_dit(Pv, Pa, Pb, cfg, inp["visual_x"], inp["audio_x"], inp["visual_context"],
                                      inp["audio_context"], inp["visual_t_mod"], inp["audio_t_mod"], inp["visual_freqs"],
                                      inp["audio_freqs"], cfg["grid_size"], 24.0, condition_scale=0.7)
    assert (fv - rv).abs().max() <= 5e-5 * max(rv.abs().max().item(), 1.0)
    assert (fa - ra).abs().max() <= 5e-5 * max(ra.abs().max().item(), 1.0)


def test_sp_split_dim0_and_gather_match_reference():
    R = ref_loader.load()
    x = torch.arange(11 * 1 * 4, dtype=torch.float32).reshape(11, 1, 4)
    for sp in (2, 3, 4, 8, 16):
        chunks = []
        for r in range(sp):
            ref, chunk_len, pad_len, total = R.functional._sp_split_tensor_dim_0(x, sp_size=sp, sp_rank=r)
            got, cl, pl, tt = O.sp_split(x, sp, r, dim=0)
            assert torch.equal(got, ref) and (cl, pl, tt) == (chunk_len, pad_len, total)
            chunks.append(got)
        assert torch.equal(O.sp_gather(chunks, pl, dim=0), x)


@pytest.mark.parametrize("seed,grid,audio_len,timestep", [(5, (2, 3, 2), 9, 37.5), (11, (1, 2, 6), 14, 999.0)])
def test_step_matches_reference(seed, grid, audio_len, timestep):
    """MOVA.inference_single_step (source lifted from pipeline_mova.py:500-609) at other seeds / geometries /
    timesteps than the committed golden."""
    import make_golden

    cfg = dict(O.TINY_STEP_CFG, grid_size=grid, audio_len=audio_len, timestep=timestep)
    arrays, _, _ = make_golden.run_reference_step(cfg, seed)
    Pv, Pa, Pb, inp = O.make_step_case(cfg, seed)
    v, a = O.inference_single_step(Pv, Pa, Pb, cfg, inp["visual_latents"], inp["audio_latents"], inp["context"],
                                   inp["timestep"])
    for got, key in ((v, "visual_output"), (a, "audio_output")):
        ref = torch.from_numpy(arrays[key])
        assert got.shape == ref.shape
        assert (got - ref).abs().max() <= 5e-5 * max(ref.abs().max().item(), 1.0)


def _load_reference_schedulers():
    """flow_match.py / flow_match_pair.py with mmengine's registry and diffusers' SchedulerMixin stubbed (neither is
    installed here; both only decorate / subclass, no arithmetic)."""
    import importlib
    import sys
    import types

    ref_loader.load()
    if "mova.registry" not in sys.modules or not hasattr(sys.modules["mova.registry"], "DIFFUSION_SCHEDULERS"):
        reg = types.ModuleType("mova.registry")

        class _Registry:
            def register_module(self, *a, **k):
                return lambda cls: cls

        reg.DIFFUSION_SCHEDULERS = _Registry()
        sys.modules["mova.registry"] = reg
    if "diffusers.schedulers.scheduling_utils" not in sys.modules:
        sch = types.ModuleType("diffusers.schedulers")
        su = types.ModuleType("diffusers.schedulers.scheduling_utils")
        su.SchedulerMixin = type("SchedulerMixin", (), {})
        sch.scheduling_utils = su
        sys.modules["diffusers.schedulers"] = sch
        sys.modules["diffusers.schedulers.scheduling_utils"] = su
    pkg = "mova.diffusion.schedulers"
    if pkg not in sys.modules:
        m = types.ModuleType(pkg)
        m.__path__ = [__import__("os").path.join(ref_loader.REFERENCE_ROOT, "mova", "diffusion", "schedulers")]
        sys.modules[pkg] = m
    return importlib.import_module(pkg + ".flow_match_pair")


@pytest.mark.parametrize("steps", [4, 50])
def test_scheduler_tables_and_update_match_reference(steps):
    """The oracle's PairScheduler / guided_update against the reference's FlowMatchPairScheduler (MOVA config: shift 5,
    extra_one_step): inference tables, the train-table sigma lookup and step_from_to."""
    fmp = _load_reference_schedulers()
    ref = fmp.FlowMatchPairScheduler(num_inference_steps=steps, num_train_timesteps=1000, shift=5.0, extra_one_step=True)
    mine = O.PairScheduler(num_inference_steps=steps)
    assert torch.equal(ref.get_pairs(), mine.get_pairs())
    assert torch.equal(ref.train_sigmas, mine.train_sigmas)
    pairs = ref.get_pairs()
    g = torch.Generator().manual_seed(1)
    sample, posi, nega = (torch.randn(2, 3, 5, generator=g) for _ in range(3))
    for i in range(steps):
        t = pairs[i, 0]
        nxt = pairs[i + 1, 0] if i + 1 < steps else None
        assert float(ref.timestep_to_sigma(t)) == float(mine.timestep_to_sigma(t))
        noise = nega + 5.0 * (posi - nega)
        want = ref.step_from_to(noise, t, nxt, sample)
        got = O.guided_update(posi, nega, sample, 5.0, float(mine.timestep_to_sigma(t)),
                              float(mine.timestep_to_sigma(nxt)) if nxt is not None else 0.0)
        assert (want - got).abs().max() <= 1e-6
