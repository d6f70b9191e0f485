"""Importance subnet on the GPU (csrc/subnet.cu through vrvq_snake_conv3_f32): against the reference fixtures and the numpy
oracle, single blocks on awkward shapes and strided views, config-2 size through shard equivalence, and the VBR forward that
feeds the map into the fused encode kernel."""
import numpy as np
import pytest
import torch

from oracle import subnet_port as sp
from tests import helpers as H
from tests.golden import gen_inputs as gi
from tests.test_subnet_cpu import IMP_ATOL, case_inputs

pytestmark = pytest.mark.gpu


def build(c, sd):
    from vrvq_b200.layers import ImportanceSubnet

    m = ImportanceSubnet(d_input=c["d_input"], d_feat=c["d_feat"], intermediate_channels=list(c["widths"]))
    m.load_state_dict(gi.torch_state_dict(sd), strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("name", list(gi.SUBNET_CASES))
def test_subnet_matches_reference_fixture_and_oracle(name):
    c = gi.SUBNET_CASES[name]
    sd, x = case_inputs(c)
    m = build(c, sd)
    y = m(torch.from_numpy(x).cuda()).cpu().numpy()
    g = H.load_golden(name)["imp_map"]
    exact = sp.importance_subnet(sd, x, dtype=np.float64)
    assert y.shape == g.shape and y.dtype == np.float32
    assert np.abs(y - g).max() <= IMP_ATOL, "vs the reference's imp_map"
    assert np.abs(y - exact).max() <= IMP_ATOL, "vs the binary64 evaluation"


@pytest.mark.parametrize("B,cin,cout,T,sig", [(2, 8, 1, 5, True), (1, 16, 130, 257, False), (3, 24, 128, 128, False), (2, 40, 7, 129, True),
                                               (0, 8, 4, 9, False), (2, 8, 4, 0, False)])
def test_single_block_shapes(B, cin, cout, T, sig):
    from vrvq_b200 import ops

    rng = np.random.Generator(np.random.PCG64(100 + cin + cout + T))
    w = (rng.normal(size=(cout, cin, 3)) / np.sqrt(3 * cin)).astype(np.float32)
    alpha = rng.uniform(0.5, 1.5, cin).astype(np.float32)
    bias = rng.normal(size=cout).astype(np.float32)
    x = rng.normal(0, 1.5, (B, cin, T)).astype(np.float32)
    pw = ops.PackedConv3(torch.from_numpy(alpha), torch.from_numpy(w), torch.from_numpy(bias), "cuda")
    y = ops.snake_conv3(pw, torch.from_numpy(x).cuda(), sigmoid=sig).cpu().numpy()
    o = sp.conv3(sp.snake(x.astype(np.float64), alpha, np.float64), w.astype(np.float64), bias.astype(np.float64))
    if sig:
        o = 1.0 / (1.0 + np.exp(-o))
    assert y.shape == (B, cout, T)
    if y.size:
        assert np.abs(y - o).max() <= 1e-5 * max(1.0, np.abs(o).max())


def test_frame_range_view_equals_padded_recompute():
    """A strided frame-range view [t0, t0+n) of a longer tensor is a valid input (include/vrvq.h: strides in elements);
    the block treats the view's ends as sequence ends (zero padding), exactly like a contiguous copy of the range."""
    from vrvq_b200 import ops

    rng = np.random.Generator(np.random.PCG64(7))
    w = (rng.normal(size=(20, 16, 3)) / 7).astype(np.float32)
    pw = ops.PackedConv3(torch.ones(16), torch.from_numpy(w), torch.zeros(20), "cuda")
    x = torch.from_numpy(rng.normal(size=(2, 16, 400)).astype(np.float32)).cuda()
    view = x[:, :, 131:301]
    assert not view.is_contiguous()
    assert torch.equal(ops.snake_conv3(pw, view), ops.snake_conv3(pw, view.contiguous()))
    # interior frames of the range equal the full-sequence result (3-tap support)
    full = ops.snake_conv3(pw, x)
    assert torch.equal(ops.snake_conv3(pw, view)[:, :, 1:-1], full[:, :, 132:300])


def test_config2_size_shards_and_sampled_oracle():
    """B=16 x T=862 (BASELINE.json configs[1]) on the shipped architecture: batch shards are bit-identical to the full call
    (the subnet shards by batch), results are deterministic, and two items agree with the exact value."""
    c = dict(gi.SUBNET_CASES["subnet_d1024"], B=16, T=862)
    sd, x = case_inputs(c)
    m = build(c, sd)
    xd = torch.from_numpy(x).cuda()
    y = m(xd)
    assert torch.equal(y, m(xd))
    for world in (2, 8):
        parts = [m(p.contiguous()) for p in xd.chunk(world, dim=0)]
        assert torch.equal(torch.cat(parts, 0), y)
    # against the binary64 evaluation: two fp32 evaluations (kernel and fp32 oracle) each carry ~2e-6 of their own
    o = sp.importance_subnet(sd, x[[0, 15]], dtype=np.float64)
    assert np.abs(y[[0, 15]].cpu().numpy() - o).max() <= IMP_ATOL
    assert float(y.min()) > 0.0 and float(y.max()) < 1.0


def test_vbr_forward_uses_the_subnet_kernels():
    """VBRResidualVectorQuantize.forward(z, feat_enc=..., level=...) (models/quantize.py:372-395): imp_map comes from the CUDA
    subnet, the mask from that map inside the fused encode kernel."""
    import vrvq_b200
    from vrvq_b200 import _lib

    Nq, D, B, T = 8, 1024, 2, 87
    c = gi.SUBNET_CASES["subnet_d1024"]
    sub_sd, feat = case_inputs(c)
    sd = gi.torch_state_dict(gi.make_state_dict(7, Nq, D))
    sd.update({"imp_subnet." + k: v for k, v in gi.torch_state_dict(sub_sd).items()})
    m = vrvq_b200.VBRResidualVectorQuantize(input_dim=D, n_codebooks=Nq, codebook_size=1024, codebook_dim=8, level_min=0.125, level_max=6.0)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    z = torch.from_numpy(gi.make_latents(8, B, D, T, 1.0)).cuda()
    n0 = _lib.launch_count
    r = m(z, n_quantizers=None, feat_enc=torch.from_numpy(feat).cuda(), level=0.5)
    assert _lib.launch_count - n0 == 7, "six subnet launches + one fused encode launch"
    imp = r["imp_map"]
    assert imp.shape == (B, 1, T)
    assert np.abs(imp.cpu().numpy() - H.load_golden("subnet_d1024")["imp_map"]).max() <= IMP_ATOL
    assert torch.equal(r["mask_imp"], vrvq_b200.generate_mask_hard(imp * 0.5 * Nq, Nq))


def test_every_tile_width_gives_identical_results(monkeypatch):
    """The launcher picks 32/64/96/128-frame tiles by wave count (csrc/subnet.cu: pick_nj); VRVQ_SUBNET_NJ forces each
    instantiation: all four must agree bit for bit (every output sums its terms in the same order)."""
    from vrvq_b200 import ops

    rng = np.random.Generator(np.random.PCG64(77))
    w = (rng.normal(size=(200, 64, 3)) / 14).astype(np.float32)
    pw = ops.PackedConv3(torch.from_numpy(rng.uniform(0.5, 1.5, 64).astype(np.float32)), torch.from_numpy(w),
                         torch.from_numpy(rng.normal(size=200).astype(np.float32)), "cuda")
    x = torch.from_numpy(rng.normal(size=(3, 64, 391)).astype(np.float32)).cuda()
    outs = []
    for nj in (1, 2, 3, 4):
        monkeypatch.setenv("VRVQ_SUBNET_NJ", str(nj))
        outs.append(ops.snake_conv3(pw, x))
    monkeypatch.delenv("VRVQ_SUBNET_NJ")
    ref = ops.snake_conv3(pw, x)
    assert all(torch.equal(o, ref) for o in outs)
    o = sp.conv3(sp.snake(x.cpu().numpy().astype(np.float64), pw.alpha.cpu().numpy(), np.float64), w.astype(np.float64),
                 pw.bias.cpu().numpy().astype(np.float64))
    assert np.abs(ref.cpu().numpy() - o).max() <= 1e-5 * np.abs(o).max()
