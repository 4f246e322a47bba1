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
