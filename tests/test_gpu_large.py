"""GPU: BASELINE.json configs 3 and 4 at (near) full size through size-independent properties plus oracle spot checks."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests import helpers as H
from tests.golden import gen_inputs as gi

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


def test_config3_nq28_against_oracle_every_item():
    """conf/base_24kbps.yml: Nq=28; B=64 x T=862 on the GPU, every item audited against the (OpenMP) oracle."""
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(71, 28, 1024))
    w = c_oracle.OracleWeights.from_state_dict(sd)
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    B, T = 64, 862
    z = torch.randn(B, 1024, T, generator=torch.Generator().manual_seed(72)).cuda()
    out = ops.rvq_encode(pw, z, None, None, None, want_z_q_is=False)
    idx = list(range(B))
    o = c_oracle.encode(w, npy(z[idx]), None, want_z_q_is=False)
    excused, skip = H.assert_codes_match(w, o, npy(out.codes[idx]))
    H.assert_close_frames(npy(out.z_q[idx]), o["z_q"], skip=skip, what="z_q")
    assert int(out.kept.sum().item()) == 28 * B * T and bool((out.mask == 1).all())
    # decode side round trip: from_codes(codes) reproduces z_q up to the straight-through rounding (quantize.py:73-75)
    zq2, _, _ = ops.from_codes(pw, out.codes, want_z_p=False)
    H.assert_close_frames(npy(zq2[idx]), npy(out.z_q[idx]), rtol=1e-5, what="from_codes(encode(z).codes) vs z_q")
    print(f"config-3 shape: {excused} audited near-tie frames of {len(idx) * T}")


def test_config4_long_form_shard_properties():
    """60 s items (T=5168): one 8-GPU shard's worth of batch (B=32 would be 677 MB; B=8 here) -- frame-split and batch-split
    launches reproduce the single launch bit-for-bit, z_q = sum of masked z_q_is (fp32 tolerance), kept counts match the mask."""
    from vrvq_b200 import ops, sharding

    sd = gi.torch_state_dict(gi.make_state_dict(81, 8, 1024))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    B, T = 8, 5168
    g = torch.Generator().manual_seed(82)
    z = torch.randn(B, 1024, T, generator=g).cuda()
    imp = torch.rand(B, 1, T, generator=g).cuda()
    full = ops.rvq_encode(pw, z, None, imp, 0.7, want_z_q_is=True)
    # (1) masked sum of the per-stage outputs, ascending stage order (quantize.py:421), equals the fused z_q within the fp32
    # tolerance (the fused z_q is one tensor-core GEMM over the kept stages, not a sum of the rounded stage outputs)
    acc = torch.zeros_like(full.z_q)
    for k in range(8):
        acc = acc + full.z_q_is[:, k] * full.mask[:, k:k + 1, :]
    H.assert_close_frames(npy(full.z_q), npy(acc), what="fused z_q vs masked sum of z_q_is")
    # (2) kept counts == column sums of the mask; mask is a prefix of ones per frame
    assert torch.equal(full.kept, full.mask.sum(dim=(0, 2)).to(torch.int64))
    assert bool((full.mask[:, 1:, :] <= full.mask[:, :-1, :]).all())
    # (3) 8-way batch x frame shards == the single launch
    kept = torch.zeros_like(full.kept)
    for segs in sharding.plan_shards(B, T, 8):
        part = sharding.encode_shard(pw, z, segs, 8, imp, 0.7, want_z_q_is=False)
        for s in segs:
            sl = (slice(s.b, s.b + 1), slice(None), slice(s.t0, s.t1))
            assert torch.equal(part.codes[sl], full.codes[sl]) and torch.equal(part.z_q[sl], full.z_q[sl])
        kept += part.kept
    assert torch.equal(kept, full.kept)
    # (4) every frame of every item against the oracle (41 k frames x 8 stages: seconds with OpenMP)
    w = c_oracle.OracleWeights.from_state_dict(sd)
    total = 0
    for b in range(B):
        o = c_oracle.encode(w, npy(z[b:b + 1]), None, npy(imp[b:b + 1]), 0.7, want_z_q_is=False)
        excused, skip = H.assert_codes_match(w, o, npy(full.codes[b:b + 1]))
        total += excused
        assert np.array_equal(npy(full.mask[b:b + 1]), o["mask"])
        H.assert_close_frames(npy(full.z_q[b:b + 1]), o["z_q"], skip=skip, what=f"z_q item {b}")
    print(f"config-4 shape: {total} audited near-tie frames of {B * T}")


def test_remask_level_sweep_matches_fused_encode():
    """scripts/inference.py:95-112: encode once, re-mask per level == encode at that level (level*Nq exactly representable)."""
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(91, 8, 1024))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    B, T = 4, 431
    g = torch.Generator().manual_seed(92)
    z = torch.randn(B, 1024, T, generator=g).cuda()
    imp = torch.rand(B, 1, T, generator=g).cuda()
    base = ops.rvq_encode(pw, z, None, imp, 1.0, want_z_q_is=True)
    for level in (0.25, 0.5, 1.0, 2.0):  # powers of two: imp*level*8 == imp*(level*8) bit-for-bit
        fused = ops.rvq_encode(pw, z, None, imp, level, want_z_q_is=False)
        zq, mask, kept = ops.remask(base.z_q_is, imp, level * 8)
        assert torch.equal(mask, fused.mask) and torch.equal(kept, fused.kept)
        H.assert_close_frames(npy(zq), npy(fused.z_q), what=f"re-masked z_q at level {level}")
