"""-m gpu: both attention schedules the library carries -- 92 (round-2 kernel, CTA pair: long key sequences) and 3
(round-1 kernel: a handful of key blocks against many queries) -- forced on EVERY shape, against the CPU oracle through ``ops.attention(variant=...)``
on ragged shapes (query / key counts that are not multiples of the 128-row tiles, odd tile counts so that the second
CTA of the last pair is empty, a single key block so that warpgroup B has nothing to do, peaked scores that force the
shared reference maximum to advance and the accumulator to be rescaled by either warpgroup).  Timing comparisons live
in benchmarks/kernels_vs_libs.py, not here.  The file sorts last so it cannot disturb the parity tests."""
import pytest
import torch

import mova_oracle as O
from util import assert_close

pytestmark = [pytest.mark.gpu]

# (B, Sq, Skv, H)
SHAPES = [(1, 128, 128, 1), (1, 100, 77, 2), (1, 256, 512, 2), (2, 300, 403, 3), (1, 403, 403, 12), (1, 1000, 512, 4),
          (1, 403, 4400, 12), (1, 129, 1300, 1), (1, 4400, 4400, 5)]


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("variant,emu", [(92, 4), (3, 4), (3, 0), (3, 8)])
def test_attention_variants_vs_oracle(variant, emu):
    import dualforce_b200 as B

    B._lib.require_device(0)
    for i, (b, sq, skv, h) in enumerate(SHAPES):
        q, k, v = _rand((b, sq, h * 128), 3 * i), _rand((b, skv, h * 128), 3 * i + 1), _rand((b, skv, h * 128), 3 * i + 2)
        got, lse = B.ops.attention(q.cuda(), k.cuda(), v.cuda(), h, return_lse=True, variant=variant, emu=emu)
        ref, ref_lse = O.attention(q.float(), k.float(), v.float(), h, return_lse=True)
        assert_close(got, ref, f"variant {variant} emu {emu} shape {(b, sq, skv, h)}", ratio=1.5e-2, fro=1e-2)
        assert (lse.cpu() - ref_lse).abs().max() <= 2e-3 * max(1.0, ref_lse.abs().max().item())


@pytest.mark.parametrize("variant", [92, 3])
def test_reference_maximum_advances_late(variant):
    """Keys sorted so that the row maximum keeps growing by more than the lazy-rescale threshold (2^8) from block to
    block: every block takes the rescale path, alternately in warpgroup A and B, and the private row sums must be
    rebased each time.  Also the reverse order (maximum in block 0: never rescaled) and a single huge outlier key."""
    import dualforce_b200 as B

    B._lib.require_device(0)
    H, Sq, Skv = 2, 200, 1100
    g = torch.Generator().manual_seed(7)
    q = torch.randn(1, Sq, H * 128, generator=g)
    k = torch.randn(1, Skv, H * 128, generator=g)
    v = torch.randn(1, Skv, H * 128, generator=g)
    qdir = torch.nn.functional.normalize(q.reshape(1, Sq, H, 128).mean(dim=1, keepdim=True), dim=-1)  # [1,1,H,128]
    ramp = torch.linspace(0.0, 60.0, Skv).reshape(1, Skv, 1, 1)
    for name, kk in (("growing", k.reshape(1, Skv, H, 128) + ramp * qdir),
                     ("shrinking", k.reshape(1, Skv, H, 128) + ramp.flip(1) * qdir),
                     ("outlier", k.reshape(1, Skv, H, 128) + (torch.arange(Skv) == 777).reshape(1, Skv, 1, 1) * 80.0 * qdir)):
        qb = (q + 3.0 * qdir.reshape(1, 1, H * 128)).to(torch.bfloat16)
        kb, vb = kk.reshape(1, Skv, H * 128).to(torch.bfloat16), v.to(torch.bfloat16)
        got, lse = B.ops.attention(qb.cuda(), kb.cuda(), vb.cuda(), H, return_lse=True, variant=variant)
        ref, ref_lse = O.attention(qb.float(), kb.float(), vb.float(), H, return_lse=True)
        assert_close(got, ref, f"variant {variant} {name} maxima", ratio=1.5e-2, fro=1e-2)
        assert (lse.cpu() - ref_lse).abs().max() <= 2e-3 * max(1.0, ref_lse.abs().max().item())


def test_long_sequence_rows_vs_oracle():
    """The shipped schedule at the real MOVA-360p key count (S_kv = 43 120 = 337 key blocks): 256 random query rows x 2
    heads against the fp32 oracle -- the lazy-rescale logic and the fp32 row sums over 337 blocks, compared value by
    value rather than through properties."""
    import dualforce_b200 as B

    B._lib.require_device(0)
    H, Skv, D = 2, 43120, 128
    g = torch.Generator().manual_seed(11)
    k = torch.randn(1, Skv, H * D, generator=g).to(torch.bfloat16)
    v = torch.randn(1, Skv, H * D, generator=g).to(torch.bfloat16)
    rows = torch.randint(0, Skv, (256,), generator=g)
    q_full = torch.randn(1, Skv, H * D, generator=g).to(torch.bfloat16)
    got, lse = B.ops.attention(q_full.cuda(), k.cuda(), v.cuda(), H, return_lse=True)
    q_rows = q_full[:, rows]
    ref, ref_lse = O.attention(q_rows.float(), k.float(), v.float(), H, return_lse=True)
    assert_close(got[:, rows.cuda()], ref, "S_kv = 43120, 256 rows x 2 heads", ratio=1.5e-2, fro=1e-2)
    assert (lse[:, :, rows.cuda()].cpu() - ref_lse).abs().max() <= 2e-3 * max(1.0, ref_lse.abs().max().item())


def test_split_kv_bridge_attention_on_device():
    """The v2a bridge shape (403 audio queries, 43 120 video keys, 12 heads) with the key sequence split into chunks
    in the batch dimension + exact LSE merge: same result as the single launch."""
    import dualforce_b200 as B

    B._lib.require_device(0)
    g = torch.Generator().manual_seed(2)
    cca = B.ConditionalCrossAttention(1536, 5120, 12)
    for prm in cca.parameters():
        prm.data = torch.randn(prm.shape, generator=g) * (0.02 if prm.dim() > 1 else 0.1)
    cca.to("cuda", torch.bfloat16)
    x = torch.randn(1, 403, 1536, generator=g).to(torch.bfloat16).cuda()
    y = torch.randn(1, 43120, 5120, generator=g).to(torch.bfloat16).cuda()
    q = cca.project_q(x, None)
    k, v = cca.project_kv(y, None)
    from dualforce_b200.modules import split_kv_attention

    plain = B.ops.attention(q, k, v, 12)
    split = split_kv_attention(q, k, v, 12, 5)
    auto = cca.attn(q, k, v)  # the attention processor picks the split factor from the shape
    assert_close(auto, plain.float().cpu(), "auto split-KV vs unsplit", ratio=1e-2, fro=6e-3)
    assert_close(split, plain.float().cpu(), "split-KV vs unsplit", ratio=1e-2, fro=6e-3)
