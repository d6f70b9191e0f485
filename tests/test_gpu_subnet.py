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
    assert _lib.launch_count - n0 == 6, "Snake pre-pass + three tensor-core blocks + the fused tail + one fused encode launch"
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


# ---- the wide blocks on the tensor cores (csrc/subnet_tc.cu through vrvq_snake_conv3_tc_f32) ---------------------------------
def _tc_case(seed, B, cin, cout, T):
    rng = np.random.Generator(np.random.PCG64(seed))
    w = (rng.normal(size=(cout, cin, 3)) / np.sqrt(3 * cin)).astype(np.float32)
    alpha = rng.uniform(0.5, 1.5, cin).astype(np.float32)
    bias = rng.normal(size=cout).astype(np.float32)
    x = rng.normal(0, 1.5, (B, cin, T)).astype(np.float32)
    return w, alpha, bias, x


@pytest.mark.parametrize("B,cin,cout,T", [(2, 64, 128, 130), (1, 96, 256, 257), (3, 64, 128, 100), (2, 128, 384, 87), (1, 64, 128, 1),
                                           (2, 64, 128, 3), (1, 160, 128, 129), (5, 64, 256, 127)])
def test_tensor_core_block_shapes(B, cin, cout, T, monkeypatch):
    """One tensor-core block (frame tiles with ragged ends, odd row pitches -> 1, 2 or 4 channel classes of tensor maps, several
    output-channel tiles) against the binary64 value and against the CUDA-core kernel of the same block."""
    from vrvq_b200 import ops

    w, alpha, bias, x = _tc_case(500 + B + cin + cout + T, B, cin, cout, T)
    pw = ops.PackedConv3(torch.from_numpy(alpha), torch.from_numpy(w), torch.from_numpy(bias), "cuda")
    assert pw.packed_tc is not None
    xd = torch.from_numpy(x).cuda()
    y = ops.snake_conv3(pw, xd).cpu().numpy()
    monkeypatch.setenv("VRVQ_SUBNET_IMPL", "cuda")
    y_cuda = ops.snake_conv3(pw, xd).cpu().numpy()
    monkeypatch.delenv("VRVQ_SUBNET_IMPL")
    o = sp.conv3(sp.snake(x.astype(np.float64), alpha, np.float64), w.astype(np.float64), bias.astype(np.float64))
    scale = max(1.0, np.abs(o).max())
    assert y.shape == (B, cout, T)
    assert np.abs(y - o).max() <= 4e-6 * scale, "3xTF32 with drained accumulators is fp32-grade"
    assert np.abs(y - y_cuda).max() <= 6e-6 * scale
    assert not np.array_equal(y, y_cuda) or T <= 3, "the two kernels sum in different orders: identical outputs mean the tensor-core path did not run"


def test_tensor_core_block_pre_and_post_activation():
    """alpha == NULL (input already activated by vrvq_snake_f32) and post_alpha (output stored through the next block's Snake)
    compose to the same chain as two plain blocks."""
    from vrvq_b200 import ops

    w0, a0, b0, x = _tc_case(901, 2, 64, 128, 301)
    w1, a1, b1, _ = _tc_case(902, 1, 128, 128, 1)
    p0 = ops.PackedConv3(torch.from_numpy(a0), torch.from_numpy(w0), torch.from_numpy(b0), "cuda")
    p1 = ops.PackedConv3(torch.from_numpy(a1), torch.from_numpy(w1), torch.from_numpy(b1), "cuda")
    xd = torch.from_numpy(x).cuda()
    plain = ops.snake_conv3(p1, ops.snake_conv3(p0, xd))
    xs = ops.snake(xd, p0.alpha)
    o64 = sp.snake(x.astype(np.float64), a0, np.float64)
    assert np.abs(xs.cpu().numpy() - o64).max() <= 2e-6 * np.abs(o64).max()
    h = ops.snake_conv3(p0, xs, pre_activated=True, post_alpha=p1.alpha)
    fused = ops.snake_conv3(p1, h, pre_activated=True)
    assert torch.equal(fused, plain), "same operations in the same order, only moved between launches"
    o = sp.conv3(sp.snake(sp.conv3(o64, w0.astype(np.float64), b0.astype(np.float64)), a1, np.float64), w1.astype(np.float64), b1.astype(np.float64))
    assert np.abs(fused.cpu().numpy() - o).max() <= 5e-6 * np.abs(o).max()


def test_tensor_core_block_on_views():
    """Channel-range and frame-range views (unit stride along T, any row pitch): the tensor maps are built per call from the strides."""
    from vrvq_b200 import ops

    w, alpha, bias, x = _tc_case(903, 2, 64, 128, 400)
    pw = ops.PackedConv3(torch.from_numpy(alpha), torch.from_numpy(w), torch.from_numpy(bias), "cuda")
    big = torch.from_numpy(np.random.Generator(np.random.PCG64(9)).normal(size=(2, 96, 517)).astype(np.float32)).cuda()
    big[:, 16:80, 57:457] = torch.from_numpy(x).cuda()
    view = big[:, 16:80, 57:457]
    assert not view.is_contiguous()
    assert torch.equal(ops.snake_conv3(pw, view), ops.snake_conv3(pw, torch.from_numpy(x).cuda()))


# ---- the narrow tail 128 -> 32 -> 8 -> 1 + sigmoid in one launch (vrvq_subnet_tail_f32) ------------------------------------------
def _tail_blocks(seed):
    from vrvq_b200 import ops

    rng = np.random.Generator(np.random.PCG64(seed))
    raw, blocks = [], []
    for cin, cout in ((128, 32), (32, 8), (8, 1)):
        w = (rng.normal(size=(cout, cin, 3)) / np.sqrt(3 * cin) * 1.5).astype(np.float32)
        alpha = rng.uniform(0.5, 1.5, cin).astype(np.float32)
        bias = rng.normal(size=cout).astype(np.float32)
        raw.append((w, alpha, bias))
        blocks.append(ops.PackedConv3(torch.from_numpy(alpha), torch.from_numpy(w), torch.from_numpy(bias), "cuda"))
    return raw, blocks


@pytest.mark.parametrize("B,T", [(2, 87), (1, 1), (3, 60), (2, 61), (1, 59), (2, 121), (4, 301), (0, 5), (2, 0)])
def test_fused_tail_matches_block_chain_and_binary64(B, T, monkeypatch):
    """Tile edges (60 output frames per CTA), sequence ends (every level has its own zero padding) and empty inputs: against the chain
    of generic block launches and against the binary64 evaluation."""
    from vrvq_b200 import ops

    raw, blocks = _tail_blocks(40 + B + T)
    x = np.random.Generator(np.random.PCG64(B * 1000 + T)).normal(0, 1.5, (B, 128, T)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    y = ops.subnet_tail(blocks, xd)
    assert y.shape == (B, 1, T)
    if B * T == 0:
        return
    chain = xd
    for i, blk in enumerate(blocks):
        chain = ops.snake_conv3(blk, chain, sigmoid=(i == 2))
    o = x.astype(np.float64)
    for w, alpha, bias in raw:
        o = sp.conv3(sp.snake(o, alpha, np.float64), w.astype(np.float64), bias.astype(np.float64))
    o = 1.0 / (1.0 + np.exp(-o))
    assert np.abs(y.cpu().numpy() - o).max() <= IMP_ATOL
    assert np.abs(y.cpu().numpy() - chain.cpu().numpy()).max() <= IMP_ATOL  # (two fp32 evaluations of the map)
    # pre-activated input (what a tensor-core producer's post_alpha stores) and a frame-range view give the same map
    y2 = ops.subnet_tail(blocks, ops.snake(xd, blocks[0].alpha), pre_activated=True)
    assert torch.equal(y2, y)
    big = torch.zeros((B, 128, T + 9), device="cuda")
    big[:, :, 4:4 + T] = xd
    assert torch.equal(ops.subnet_tail(blocks, big[:, :, 4:4 + T]), y)


def test_chain_with_and_without_the_fused_launches_agree(monkeypatch):
    """The shipped architecture with VRVQ_SUBNET_IMPL=cuda (six generic launches on the CUDA cores) against the default chain
    (tensor-core blocks + fused tail): two fp32-grade evaluations of the same map."""
    c = dict(gi.SUBNET_CASES["subnet_d1024"], B=3, T=301)
    sd, x = case_inputs(c)
    m = build(c, sd)
    xd = torch.from_numpy(x).cuda()
    y = m(xd)
    monkeypatch.setenv("VRVQ_SUBNET_IMPL", "cuda")
    y_cuda = m(xd)
    monkeypatch.delenv("VRVQ_SUBNET_IMPL")
    assert not torch.equal(y, y_cuda)
    assert float((y - y_cuda).abs().max()) <= 2 * IMP_ATOL  # each is within IMP_ATOL of the reference's map (fixture tests)
