"""-m gpu: BASELINE.json configs[0] on the CUDA path -- MOVA-360p widths (5120 / 40 heads / ffn 13824 and 1536 / 12 /
8960), 2 + 2 blocks, 2 bridge layers, L_v = 4400, L_a = 36, 512 text tokens -- against sampled outputs of the REFERENCE
itself (fp32 weights, fp32 CPU; tests/golden/reduced_360p_samples.npz)."""
import pytest
import torch

from test_oracle_golden import load_reduced_case
from util import build_towers, metrics, to_dev

pytestmark = [pytest.mark.gpu]


def test_reduced_360p_forward_vs_reference_samples():
    cfg, Pv, Pa, Pb, inp, gold, idx = load_reduced_case()
    vis, aud, bridge, pipe = build_towers(cfg, Pv, Pa, Pb)  # weights rounded to bf16 on the device
    d = to_dev(inp)
    fv, fa = pipe.forward_dual_tower_dit(vis, d["visual_x"], d["audio_x"], d["visual_context"], d["audio_context"],
                                         d["visual_t_mod"], d["audio_t_mod"], d["visual_freqs"], d["audio_freqs"],
                                         cfg["grid_size"], cfg["video_fps"])
    for name, t, x in (("visual", fv, inp["visual_x"]), ("audio", fa, inp["audio_x"])):
        got = t.float().cpu().reshape(-1)[idx[name]]
        m = metrics(got, gold[f"{name}_samples"])
        # bf16 weights + bf16 activations against the reference's fp32 run, two layers deep
        assert m["finite"] and m["ratio"] <= 4e-2 and m["rel_fro"] <= 2e-2 and m["cos"] >= 0.999, (name, m)
        delta = got - x.reshape(-1)[idx[name]].to(torch.bfloat16).float()
        md = metrics(delta, gold[f"{name}_delta_samples"])
        assert md["cos"] >= 0.995, (name, "residual delta", md)
