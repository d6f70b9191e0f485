"""The tensor-core encode kernel (csrc/rvq_encode_tc.cu) against the CUDA-core kernel (csrc/rvq_encode.cu) and the oracle.

Both kernels sit behind the same C-ABI entry point (vrvq_rvq_encode_f32); VRVQ_ENCODE_IMPL=cuda forces the CUDA-core
kernel.  They decide codes with the same exact fp32 search, so codes may differ only through the rounding of z_e
(projected-residual GEMM vs. residual FMA chains): audited near-ties, as for any two conv implementations.
"""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests import helpers as H
from tests.golden import gen_inputs as gi

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


def run_impl(impl, fn):
    old = os.environ.get("VRVQ_ENCODE_IMPL")
    if impl is None:
        os.environ.pop("VRVQ_ENCODE_IMPL", None)
    else:
        os.environ["VRVQ_ENCODE_IMPL"] = impl
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop("VRVQ_ENCODE_IMPL", None)
        else:
            os.environ["VRVQ_ENCODE_IMPL"] = old


def test_tc_kernel_is_the_default_path():
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(5, 8, 1024))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    tc = run_impl(None, lambda: ops.encode_launch_info(pw, 16, 862, 8, "cuda", z_q_is=True))
    cc = run_impl("cuda", lambda: ops.encode_launch_info(pw, 16, 862, 8, "cuda", z_q_is=True))
    assert tc["kernel"] == "tc" and tc["block"] == 512 and cc["kernel"] == "cuda", (tc, cc)
    assert tc["grid"] == 144  # with z_q_is a tile's time is its stores: 9 tiles of 96 frames per item, one wave on 148 SMs
    # without z_q_is a tile costs the same whatever its length: the fewest waves win (config-4 shard: 9 waves of 128-frame tiles)
    nz = run_impl(None, lambda: ops.encode_launch_info(pw, 32, 5168, 8, "cuda"))
    assert nz["kernel"] == "tc" and nz["grid"] == 148
    small = run_impl(None, lambda: ops.encode_launch_info(pw, 1, 87, 8, "cuda"))
    assert small["kernel"] == "cuda", "calls that fit one wave of the CUDA-core kernel stay on it (lower latency)"
    forced = run_impl("tc", lambda: ops.encode_launch_info(pw, 1, 87, 8, "cuda"))
    assert forced["kernel"] == "tc"


@pytest.mark.parametrize("D,Nq,B,T,n_run,vbr", [
    (1024, 8, 3, 87, 8, True),      # configs[0]-like, ragged
    (1024, 8, 2, 431, 8, True),     # several tiles per item, odd T (no sector shift: odd row pitch)
    (1024, 8, 2, 250, 5, False),    # CBR early exit (quantize.py:183-184)
    (512, 4, 2, 130, 4, True),
    (256, 3, 5, 33, 3, True),
    (1024, 8, 1, 1, 8, True),
    (1024, 1, 2, 64, 1, False),
])
def test_tc_matches_cuda_core_kernel_and_oracle(D, Nq, B, T, n_run, vbr):
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(100 + Nq, Nq, D))
    w = c_oracle.OracleWeights.from_state_dict(sd)
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    z_np = gi.make_latents(200 + T, B, D, T, 1.0)
    imp_np = gi.make_imp_map(300 + T, B, T) if vbr else None
    z = torch.from_numpy(z_np).cuda()
    imp = torch.from_numpy(imp_np).cuda() if vbr else None
    level = 0.6 if vbr else None

    def call():
        return ops.rvq_encode(pw, z, n_run, imp, level, want_z_q_is=True, want_loss_pf=True)

    a = run_impl("tc", call)
    c = run_impl("cuda", call)
    o = c_oracle.encode(w, z_np, n_run if not vbr else None, imp_np, level, want_z_q_is=True)
    excused, skip = H.assert_codes_match(w, o, npy(a.codes))
    assert np.array_equal(npy(a.mask), npy(c.mask)) and np.array_equal(npy(a.mask), o["mask"])
    assert np.array_equal(npy(a.kept), o["kept"])
    H.assert_close_frames(npy(a.z_q), o["z_q"], skip=skip, what="z_q vs oracle")
    H.assert_close_frames(npy(a.latents), o["latents"], skip=skip, what="latents vs oracle")
    H.assert_close_frames(npy(a.z_q_is).reshape(B, -1, T), o["z_q_is"].reshape(B, -1, T), skip=skip, what="z_q_is vs oracle")
    same = (npy(a.codes) == npy(c.codes)).all(axis=1)  # frames on which the two kernels agree on every stage
    # both kernels round z_e differently, so they may part at an fp32 near-tie; every run so far: 0 frames
    assert (~same).sum() <= max(1, int(1e-3 * same.size)), f"tensor-core and CUDA-core kernels disagree on {(~same).sum()} of {same.size} frames"
    sk = ~same
    H.assert_close_frames(npy(a.z_q), npy(c.z_q), skip=sk, what="z_q: tensor-core vs CUDA-core kernel")
    H.assert_close_frames(npy(a.z_q_is).reshape(B, -1, T), npy(c.z_q_is).reshape(B, -1, T), skip=sk, what="z_q_is: tc vs cuda-core")
    np.testing.assert_allclose(npy(a.loss_pf)[~sk[:, None, :].repeat(n_run, 1)], npy(c.loss_pf)[~sk[:, None, :].repeat(n_run, 1)], rtol=2e-4, atol=1e-7)


@pytest.mark.parametrize("t_lo,t_hi", [(0, 200), (1, 201), (3, 117), (40, 296)])
def test_tc_frame_views_any_alignment(t_lo, t_hi):
    """Frame-range views of larger tensors (the multi-GPU / chunked shard): every start alignment of the
    outputs (the sector-shift classes of the epilogue) reproduces the full call bit-for-bit, and nothing is written outside."""
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(77, 8, 1024))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    B, T = 2, 300
    g = torch.Generator().manual_seed(78)
    z = torch.randn(B, 1024, T, generator=g).cuda()
    imp = torch.rand(B, 1, T, generator=g).cuda()
    os.environ["VRVQ_ENCODE_IMPL"] = "tc"  # small shapes default to the CUDA-core kernel; this test is about the tensor-core one
    full = ops.rvq_encode(pw, z, None, imp, 0.7, want_z_q_is=True)
    n = t_hi - t_lo
    out = ops.EncodeOutputs(B, 1024, T, 8, "cuda", z_q=True, z_q_is=True, latents=True, mask=True)
    for t in (out.z_q, out.z_q_is, out.latents, out.mask):
        t.fill_(-7.0)
    out.codes.fill_(-7)
    view = ops.EncodeOutputs.__new__(ops.EncodeOutputs)
    view.codes, view.z_q, view.z_q_is = out.codes[:, :, t_lo:t_hi], out.z_q[:, :, t_lo:t_hi], out.z_q_is[:, :, :, t_lo:t_hi]
    view.latents, view.mask, view.loss_pf = out.latents[:, :, t_lo:t_hi], out.mask[:, :, t_lo:t_hi], None
    view.accum, view.n_run, view.frames = out.accum, 8, B * n
    ops.rvq_encode_into(pw, z[:, :, t_lo:t_hi], view, 8, imp[:, :, t_lo:t_hi], 0.7)
    torch.cuda.synchronize()
    for name in ("codes", "z_q", "z_q_is", "latents", "mask"):
        got, ref = getattr(out, name), getattr(full, name)
        assert torch.equal(got[..., t_lo:t_hi], ref[..., t_lo:t_hi]), f"{name}: view [{t_lo},{t_hi}) differs from the full call"
        outside = torch.cat([got[..., :t_lo].reshape(-1), got[..., t_hi:].reshape(-1)])
        assert bool((outside == -7).all()), f"{name}: wrote outside the view"
    os.environ.pop("VRVQ_ENCODE_IMPL", None)


def test_tc_optional_outputs_and_nan_latent():
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(55, 8, 1024))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    B, T = 2, 140
    g = torch.Generator().manual_seed(56)
    z = torch.randn(B, 1024, T, generator=g).cuda()
    imp = torch.rand(B, 1, T, generator=g).cuda()
    os.environ["VRVQ_ENCODE_IMPL"] = "tc"
    ref = ops.rvq_encode(pw, z, None, imp, 0.5, want_z_q_is=True)
    only_codes = ops.EncodeOutputs(B, 1024, T, 8, "cuda", z_q=False, z_q_is=False, latents=False, mask=False)
    ops.rvq_encode_into(pw, z, only_codes, 8, imp, 0.5)
    assert torch.equal(only_codes.codes, ref.codes) and torch.equal(only_codes.kept, ref.kept)
    no_zq = ops.EncodeOutputs(B, 1024, T, 8, "cuda", z_q=False, z_q_is=True, latents=True, mask=True)
    ops.rvq_encode_into(pw, z, no_zq, 8, imp, 0.5)
    assert torch.equal(no_zq.z_q_is, ref.z_q_is) and torch.equal(no_zq.codes, ref.codes)
    # an all-zero frame (degenerate search: every score ties) and a NaN frame must not disturb their neighbours
    z2 = z.clone()
    z2[0, :, 5] = 0.0
    z2[1, :, 9] = float("nan")
    r2 = ops.rvq_encode(pw, z2, None, imp, 0.5, want_z_q_is=False)
    keep = torch.ones(B, T, dtype=torch.bool, device="cuda")
    keep[0, 5] = False
    keep[1, 9] = False
    assert torch.equal(r2.codes.permute(0, 2, 1)[keep], ref.codes.permute(0, 2, 1)[keep])
    assert torch.equal(r2.z_q.permute(0, 2, 1)[keep], ref.z_q.permute(0, 2, 1)[keep])
    zf = c_oracle.encode(c_oracle.OracleWeights.from_state_dict(sd), npy(z2[0:1, :, 5:6]), None, npy(imp[0:1, :, 5:6]), 0.5)
    os.environ.pop("VRVQ_ENCODE_IMPL", None)
    assert np.array_equal(npy(r2.codes[0:1, :, 5:6]), zf["codes"]), "all-zero latent: exact fallback scan must match the oracle"


@pytest.mark.parametrize("D,Nq,n,B,T,masked", [(1024, 8, 8, 2, 131, False), (1024, 8, 8, 2, 300, True), (1024, 8, 3, 1, 87, False),
                                              (512, 4, 4, 3, 40, True), (256, 3, 2, 2, 1, False)])
def test_from_codes_on_the_tensor_core_path(D, Nq, n, B, T, masked):
    """vrvq_from_codes_f32 (quantize.py:217-249) for <= 8 codebooks runs the encode kernel's gather + out_proj GEMMs (FC
    instantiation): against the CUDA-core decode kernel (bit-exact vs the oracle) -- z_p identical, z_q / z_q_is within the GEMM
    accumulation tolerance, arbitrary 0/1 masks, frame-range views, out-of-range codes reported."""
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(600 + Nq, Nq, D))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    g = torch.Generator().manual_seed(601 + T)
    codes = torch.randint(0, 1024, (B, n, T), generator=g).cuda()
    mask = (torch.rand(B, n, T, generator=g) < 0.6).float().cuda() if masked else None

    def call(c=codes, m=mask):
        return ops.from_codes(pw, c, m, want_z_q_is=True, want_z_p=True)

    zq, zp, zqis = run_impl("tc", call)
    zq_c, zp_c, zqis_c = run_impl("cuda", call)
    assert torch.equal(zp, zp_c), "z_p = the gathered codebook rows"
    H.assert_close_frames(npy(zq), npy(zq_c), rtol=5e-6, what="from_codes z_q: tensor cores vs CUDA cores")
    H.assert_close_frames(npy(zqis).reshape(B, -1, T), npy(zqis_c).reshape(B, -1, T), rtol=5e-6, what="from_codes z_q_is")
    if T >= 40:  # a frame-range view reproduces the full call bit for bit
        sl = slice(5, T - 3)
        zq_v, zp_v, _ = run_impl("tc", lambda: call(codes[:, :, sl], None if mask is None else mask[:, :, sl]))
        assert torch.equal(zq_v, zq[:, :, sl]) and torch.equal(zp_v, zp[:, :, sl])
    bad = codes.clone()
    bad[0, 0, 0] = 1024
    with pytest.raises(IndexError):
        ops.from_codes(pw, bad)


def test_random_shapes_tensor_core_vs_cuda_core():
    """Seeded random sweep (scripts/stress_tc.py): D, Nq, n_run, B, T (tile-boundary values), VBR/CBR, with/without z_q_is --
    tensor-core kernel vs CUDA-core kernel: masks and kept counts identical, >= 97 % of the frames with identical codes,
    z_q / z_q_is / latents within 1e-5 on those frames."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("stress_tc", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "stress_tc.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    worst = mod.run_cases(11, 14, verbose=False)
    assert worst <= 1e-5


@pytest.mark.parametrize("D,Nq,B,T,n_run,vbr", [
    (1024, 28, 2, 200, 28, True),    # conf/base_24kbps.yml: four stage groups, masks over 28 stages, several tiles per item
    (1024, 28, 3, 131, 12, False),   # CBR early exit inside the second group (quantize.py:183-184)
    (1024, 28, 2, 64, 3, False),     # early exit inside the first group: one pass of the grouped kernel
    (512, 9, 2, 130, 9, True),       # class defaults of ResidualVectorQuantize: one stage in the second group
    (256, 17, 4, 33, 17, True),      # three groups, ragged single tiles
    (1024, 32, 1, 87, 32, True),     # the largest supported model, odd T (four channel classes of the latent tensor maps)
    (1024, 16, 2, 1, 16, False),     # T = 1
])
def test_grouped_tensor_core_kernel_matches_cuda_core_kernel_and_oracle(D, Nq, B, T, n_run, vbr):
    """Models with more than 8 codebooks run the tensor-core kernel in groups of 8 stages (no z_q_is): per-group in_proj,
    cross-group corrections as virtual-channel chunks, final z_q from codes.  Against the oracle (audited codes, exact masks and
    kept counts, z_q / latents within 1e-5) and the CUDA-core kernel."""
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(300 + Nq + D, Nq, D))
    w = c_oracle.OracleWeights.from_state_dict(sd)
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    assert run_impl("tc", lambda: ops.encode_launch_info(pw, B, T, n_run, "cuda"))["kernel"] == "tc"
    assert run_impl("tc", lambda: ops.encode_launch_info(pw, B, T, n_run, "cuda", z_q_is=True))["kernel"] == "cuda", "z_q_is needs the CUDA-core kernel here"
    z_np = gi.make_latents(400 + T, B, D, T, 1.0)
    imp_np = gi.make_imp_map(500 + T, B, T) if vbr else None
    z = torch.from_numpy(z_np).cuda()
    imp = torch.from_numpy(imp_np).cuda() if vbr else None
    level = 0.45 if vbr else None

    def call():
        return ops.rvq_encode(pw, z, n_run, imp, level, want_z_q_is=False, want_loss_pf=True)

    a = run_impl("tc", call)
    c = run_impl("cuda", call)
    o = c_oracle.encode(w, z_np, n_run if not vbr else None, imp_np, level, want_z_q_is=False)
    excused, skip = H.assert_codes_match(w, o, npy(a.codes))
    assert np.array_equal(npy(a.mask), o["mask"]) and np.array_equal(npy(a.mask), npy(c.mask))
    assert np.array_equal(npy(a.kept), o["kept"])
    H.assert_close_frames(npy(a.z_q), o["z_q"], skip=skip, what="z_q vs oracle")
    H.assert_close_frames(npy(a.latents), o["latents"], skip=skip, what="latents vs oracle")
    same = (npy(a.codes) == npy(c.codes)).all(axis=1)
    assert (~same).sum() <= max(1, int(1e-3 * same.size))
    H.assert_close_frames(npy(a.z_q), npy(c.z_q), skip=~same, what="z_q: grouped tensor-core vs CUDA-core kernel")
    if excused == 0:
        np.testing.assert_allclose(npy(a.loss_pf), o["loss_pf"], rtol=2e-4, atol=1e-7)
        assert a.loss_sum.item() == pytest.approx(o["loss_masked_sum"], rel=1e-5)


def test_grouped_kernel_views_and_many_tiles_per_cta():
    """Config-3 shape at reduced batch: several tiles x 4 groups per CTA (ring positions are running totals across passes of
    different length), and a frame-range view reproduces the full call bit for bit."""
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(71, 28, 1024))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    B, T = 48, 431
    z = torch.randn(B, 1024, T, generator=torch.Generator().manual_seed(5)).cuda()
    full = ops.rvq_encode(pw, z, None, None, None)
    info = ops.encode_launch_info(pw, B, T, 28, "cuda")
    assert info["kernel"] == "tc" and B * -(-T // 120) > info["grid"], "more tiles than CTAs"
    part = ops.rvq_encode(pw, z[:, :, 64:300], None, None, None)
    assert torch.equal(part.codes, full.codes[:, :, 64:300]) and torch.equal(part.z_q, full.z_q[:, :, 64:300])
    w = c_oracle.OracleWeights.from_state_dict(sd)
    idx = [0, 23, 47]
    o = c_oracle.encode(w, npy(z[idx]), None, want_z_q_is=False)
    excused, skip = H.assert_codes_match(w, o, npy(full.codes[idx]))
    H.assert_close_frames(npy(full.z_q[idx]), o["z_q"], skip=skip, what="z_q")


def test_codebook_rows_that_do_not_normalise_to_unit_vectors():
    """F.normalize leaves a row with norm < 1e-12 at ~0, so its distance e2 - 0 + c2 ~ 1 beats every unit row whenever the
    best cosine is below 0.5 -- but its TF32 score is 0, which the margin filter would drop.  Such rows are listed in the blob
    (section SPC) and always re-scored exactly: the tensor-core kernel must pick them exactly where the oracle does."""
    from vrvq_b200 import ops

    D, Nq, B, T = 256, 3, 2, 150
    raw = gi.make_state_dict(21, Nq, D)
    rng = np.random.Generator(np.random.PCG64(5))
    for s in (0, 2):  # codebooks crowded into a cone, so that most latents see a best cosine < 0.5 ...
        cb = np.zeros((1024, 8), np.float32)
        cb[:, 0] = 1.0
        cb += 0.15 * rng.normal(size=cb.shape).astype(np.float32)
        cb[7] = 0.0          # ... and prefer the all-zero row
        cb[900] = 1e-20      # ... or the one whose norm is below F.normalize's eps
        raw[f"quantizers.{s}.codebook.weight"] = cb
    sd = gi.torch_state_dict(raw)
    w = c_oracle.OracleWeights.from_state_dict(sd)
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    hdr = pw.host_blob[:16].view(np.int32)
    from tests.test_abi_cpu import _tc_layout
    spc = pw.host_blob[int(hdr[6]) + _tc_layout(Nq, D)["off_spc"]:][:16 * Nq].view(np.int32).reshape(Nq, 16)
    assert spc[0, 0] == 2 and set(spc[0, 1:3]) == {7, 900} and spc[1, 0] == 0 and spc[2, 0] == 2
    z_np = gi.make_latents(22, B, D, T, 1.0)
    o = c_oracle.encode(w, z_np, None, None, None, want_z_q_is=True)
    assert np.isin(o["codes"][:, 0], (7, 900)).mean() > 0.2, "the case must exercise the degenerate rows"
    for impl in ("tc", "cuda"):
        out = run_impl(impl, lambda: ops.rvq_encode(pw, torch.from_numpy(z_np).cuda(), None, None, None, want_z_q_is=True))
        # (the two degenerate rows sit at distance 1 +- 1e-7 from EVERY frame: which of the two wins is decided by the last bit of
        # e2, so audited near-ties between codes 7 and 900 are expected here -- what must not happen is a frame that prefers a
        # unit row although a degenerate one is closer)
        excused, skip = H.assert_codes_match(w, o, npy(out.codes), max_excused_frac=0.1, what=f"codes ({impl})")
        differ = (npy(out.codes) != o["codes"])
        assert np.isin(npy(out.codes)[differ], (7, 900)).all() or excused == 0
        H.assert_close_frames(npy(out.z_q), o["z_q"], skip=skip, what=f"z_q ({impl})")


@pytest.mark.parametrize("Nq,n_run,vbr", [(8, 8, True), (8, 2, False), (8, 1, False), (9, 9, True), (17, 10, False)])
@pytest.mark.parametrize("T", [127, 128, 129, 257, 384])
def test_tiles_without_z_q_is_have_no_halo(Nq, n_run, vbr, T):
    """Without z_q_is a tile is 128 own frames (no 8-row halo: it only serves the shifted z_q_is stores), the frame threads hand the
    corrections of the later stages to warps 4-7 and the un-normalised codebook is staged in shared memory: tile boundaries at
    multiples of 128 frames, one / two stages (nothing for warps 4-7 to correct), VBR masks, and the grouped instantiation."""
    from vrvq_b200 import ops

    D, B = 1024, 2
    sd = gi.torch_state_dict(gi.make_state_dict(400 + Nq, Nq, D))
    w = c_oracle.OracleWeights.from_state_dict(sd)
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    z_np = gi.make_latents(500 + T, B, D, T, 1.0)
    imp_np = gi.make_imp_map(600 + T, B, T) if vbr else None
    z = torch.from_numpy(z_np).cuda()
    imp = torch.from_numpy(imp_np).cuda() if vbr else None
    level = 0.7 if vbr else None
    # (a call this small would be cut into shorter tiles to use more SMs: the profiling knob pins the 128-frame tiles of large calls)
    os.environ["VRVQ_DEBUG_TILE_FRAMES"] = "128"
    try:
        a = run_impl("tc", lambda: ops.rvq_encode(pw, z, n_run, imp, level, want_z_q_is=False, want_loss_pf=True))
        info = run_impl("tc", lambda: ops.encode_launch_info(pw, B, T, Nq, "cuda"))
    finally:
        os.environ.pop("VRVQ_DEBUG_TILE_FRAMES", None)
    flat = os.environ.get("VRVQ_FLAT_TILES") == "1" and T >= 128  # (forced flat tiling: tiles of the flattened frame sequence)
    assert info["kernel"] == "tc" and info["grid"] == ((B * T + 127) // 128 if flat else B * ((T + 127) // 128)), info  # ceil(T / 128) tiles per item
    o = c_oracle.encode(w, z_np, n_run if not vbr else None, imp_np, level, want_z_q_is=False)
    excused, skip = H.assert_codes_match(w, o, npy(a.codes))
    assert np.array_equal(npy(a.mask), o["mask"]) and np.array_equal(npy(a.kept), o["kept"])
    H.assert_close_frames(npy(a.z_q), o["z_q"], skip=skip, what="z_q vs oracle")
    H.assert_close_frames(npy(a.latents), o["latents"], skip=skip, what="latents vs oracle")


@pytest.mark.parametrize("Nq,n_run,vbr", [(8, 8, True), (8, 3, False), (9, 9, True), (28, 28, True)])
@pytest.mark.parametrize("B,T", [(3, 129), (5, 200), (4, 333), (2, 128)])
def test_flat_tiles_spanning_two_items(Nq, n_run, vbr, B, T):
    """Flat tiling (taken by itself when it saves a wave, e.g. config 3: 431 instead of 448 tiles; forced here): a tile is 128
    consecutive frames of the flattened (item, frame) sequence and may span two items -- two TMA boxes per chunk, per-row item for
    the importance map, the per-item level, codes, latents, mask and z_q."""
    from vrvq_b200 import ops

    D = 1024
    sd = gi.torch_state_dict(gi.make_state_dict(700 + Nq, Nq, D))
    w = c_oracle.OracleWeights.from_state_dict(sd)
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    z_np = gi.make_latents(800 + T, B, D, T, 1.0)
    imp_np = gi.make_imp_map(900 + T, B, T) if vbr else None
    z = torch.from_numpy(z_np).cuda()
    imp = torch.from_numpy(imp_np).cuda() if vbr else None
    level = 0.8 if vbr else None
    os.environ["VRVQ_FLAT_TILES"] = "1"
    try:
        a = run_impl("tc", lambda: ops.rvq_encode(pw, z, n_run, imp, level, want_z_q_is=False, want_loss_pf=True))
        info = run_impl("tc", lambda: ops.encode_launch_info(pw, B, T, Nq, "cuda"))
        # a frame-range view of a wider tensor (rows 8-byte aligned only): same result bit for bit
        zw = torch.zeros(B, D, T + 6, device="cuda")
        zw[:, :, 2:T + 2] = z
        v = run_impl("tc", lambda: ops.rvq_encode(pw, zw[:, :, 2:T + 2], n_run, imp, level, want_z_q_is=False))
    finally:
        os.environ.pop("VRVQ_FLAT_TILES", None)
    b = run_impl("tc", lambda: ops.rvq_encode(pw, z, n_run, imp, level, want_z_q_is=False, want_loss_pf=True))  # per-item tiles
    assert info["grid"] == (B * T + 127) // 128, info
    o = c_oracle.encode(w, z_np, n_run if not vbr else None, imp_np, level, want_z_q_is=False)
    excused, skip = H.assert_codes_match(w, o, npy(a.codes))
    assert np.array_equal(npy(a.mask), o["mask"]) and np.array_equal(npy(a.kept), o["kept"])
    H.assert_close_frames(npy(a.z_q), o["z_q"], skip=skip, what="z_q vs oracle")
    H.assert_close_frames(npy(a.latents), o["latents"], skip=skip, what="latents vs oracle")
    # frames are independent and every lane's arithmetic order is fixed: the tiling cannot change a bit
    assert torch.equal(a.codes, b.codes) and torch.equal(a.z_q, b.z_q) and torch.equal(a.latents, b.latents) and torch.equal(a.loss_pf, b.loss_pf)
    assert torch.equal(a.codes, v.codes) and torch.equal(a.z_q, v.z_q)
