"""GPU (B200): the CUDA path, called through the reference-shaped Python API and the C ABI underneath, against
(1) the committed fixtures produced by the real reference and (2) the C oracle on seeded inputs.

Bar (BASELINE.json north_star): codes and masks bit-exact except audited fp32 near-ties (first-divergence audit,
NEAR_TIE_EPS in tests/helpers.py); z_q / z_q_is / latents within 1e-5 relative per frame.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests import helpers as H
from tests.golden import gen_inputs as gi

pytestmark = pytest.mark.gpu

VBR = [n for n, c in gi.CASES.items() if c["kind"] == "vbr"]
CBR = [n for n, c in gi.CASES.items() if c["kind"] == "cbr"]


def build_module(case, device="cuda"):
    import vrvq_b200

    sd = gi.torch_state_dict(gi.make_state_dict(case["seed"], case["Nq"], case["D"], case["K"]))
    if case["kind"] == "vbr":
        m = vrvq_b200.VBRResidualVectorQuantize(input_dim=case["D"], n_codebooks=case["Nq"], codebook_size=case["K"], codebook_dim=8,
                                                level_min=0.125, level_max=6.0, imp2mask_alpha=2.0)
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith("imp_subnet.") for k in missing)
    else:
        m = vrvq_b200.ResidualVectorQuantize(input_dim=case["D"], n_codebooks=case["Nq"], codebook_size=case["K"], codebook_dim=8)
        m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


def npy(t):
    return t.detach().cpu().numpy()


@pytest.fixture(params=["cuda", "tc"])
def impl(request, monkeypatch):
    """Run a test once per fused encode kernel: `cuda` = rvq_encode_kernel (CUDA cores), `tc` = rvq_encode_tc_kernel
    (tcgen05, the production kernel for every call larger than one wave).  The fixtures are small, so without the override
    the size heuristic would send all of them to the CUDA-core kernel."""
    monkeypatch.setenv("VRVQ_ENCODE_IMPL", request.param)
    return request.param


def assert_kernel(impl, m_or_pw, B, T, n_run, z_q_is=False):
    """The kernel that actually serves the call is the one the test is named after."""
    from vrvq_b200 import ops

    pw = m_or_pw if isinstance(m_or_pw, ops.PackedWeights) else m_or_pw.packed_weights(torch.device("cuda", torch.cuda.current_device()))
    info = ops.encode_launch_info(pw, B, T, n_run, "cuda", z_q_is=z_q_is)
    if impl == "tc" and not ops.tc_kernel_available(pw, n_run, z_q_is=z_q_is):
        pytest.skip("no tensor-core kernel for this shape (the call is served by the CUDA-core kernel, covered by the other parameter)")
    assert info["kernel"] == impl, (impl, info)


def test_extension_is_loaded():
    """The driver records which .so the test process loaded; make sure it is ours and that nothing falls back."""
    from vrvq_b200 import _lib

    _lib.lib()
    with open("/proc/self/maps") as f:
        assert "libvrvq.so" in f.read()


@pytest.mark.parametrize("name", VBR)
def test_vbr_against_reference_fixture(name, impl):
    import vrvq_b200

    case, g = gi.CASES[name], H.load_golden(name)
    m = build_module(case)
    assert_kernel(impl, m, case["B"], case["T"], case["Nq"])
    w = H.oracle_weights_for(case)
    z_np = H.latents_for(case)
    z = torch.from_numpy(z_np).cuda()
    imp_np = gi.make_imp_map(case["imp_seed"], case["B"], case["T"])
    imp = torch.from_numpy(imp_np).cuda()
    B, T, Nq = case["B"], case["T"], case["Nq"]
    for li, level in enumerate(case["levels"]):
        lv = level if not isinstance(level, list) else torch.tensor(level, dtype=torch.float32, device="cuda").view(-1, 1, 1)
        r = m(z, n_quantizers=None, feat_enc=None, level=lv, imp_map=imp)
        o = c_oracle.encode(w, z_np, None, imp_np, level)
        assert np.array_equal(o["codes"], g["codes"])
        excused, skip = H.assert_codes_match(w, o, npy(r["codes"]))
        assert np.array_equal(npy(r["mask_imp"]), g[f"mask_{li}"]), "mask must be bit-identical to the reference"
        H.assert_close_frames(npy(r["z_q"]), g[f"z_q_{li}"], skip=skip, what="z_q vs reference")
        H.assert_close_frames(npy(r["latents"]), g["latents"], skip=skip, what="latents vs reference")
        H.assert_close_frames(npy(r["z_q_is"])[:, :, ::16, :].reshape(B, -1, T), g["z_q_is_sub"].reshape(B, -1, T), skip=skip, what="z_q_is")
        H.assert_close_frames(npy(r["z_q_is"]).reshape(B, -1, T), o["z_q_is"].reshape(B, -1, T), skip=skip, what="z_q_is vs oracle")
        if excused == 0:
            assert r["commitment_loss"].item() == pytest.approx(float(g[f"commitment_loss_{li}"]), rel=1e-5)
            assert r["codebook_loss"].item() == pytest.approx(float(g[f"codebook_loss_{li}"]), rel=1e-5)
        assert r["imp_map"] is imp and r["codes"].dtype == torch.int64 and r["mask_imp"].dtype == torch.float32
        bpf = vrvq_b200.cal_bpf_from_mask(r["mask_imp"], [10] * Nq)
        assert bpf == pytest.approx(float(g[f"bpf_{li}"]), rel=1e-6)
        assert np.array_equal(npy(r["kept_frames"]), o["kept"])
        # the README re-mask recipe through the mirrored utilities (README.md:75-79)
        lvs = (lv * Nq) if isinstance(lv, torch.Tensor) else level * Nq
        mh = vrvq_b200.generate_mask_hard(imp * lvs, nq=Nq)
        assert torch.equal(mh, r["mask_imp"])


@pytest.mark.parametrize("name", CBR)
def test_cbr_against_reference_fixture(name, impl):
    case, g = gi.CASES[name], H.load_golden(name)
    m = build_module(case)
    assert_kernel(impl, m, case["B"], case["T"], case["Nq"])
    w = H.oracle_weights_for(case)
    z_np = H.latents_for(case)
    z = torch.from_numpy(z_np).cuda()
    for qi, nq in enumerate(case["n_quantizers"]):
        r = m(z, n_quantizers=nq)
        o = c_oracle.encode(w, z_np, nq, want_z_q_is=False)
        excused, skip = H.assert_codes_match(w, o, npy(r["codes"]))
        if excused == 0:
            assert np.array_equal(npy(r["codes"]), g[f"codes_{qi}"])
            assert r["commitment_loss"].item() == pytest.approx(float(g[f"commitment_loss_{qi}"]), rel=1e-5)
        H.assert_close_frames(npy(r["z_q"]), g[f"z_q_{qi}"], skip=skip, what="z_q vs reference")
        H.assert_close_frames(npy(r["latents"]), g[f"latents_{qi}"], skip=skip, what="latents vs reference")
        if nq is None:
            codes = torch.from_numpy(g[f"codes_{qi}"]).cuda()
            zq, zp, _, zqis = m.from_codes(codes, return_z_q_is=True)
            H.assert_close_frames(npy(zq), g["from_codes_z_q"], what="from_codes z_q")
            assert np.array_equal(npy(zp), g["from_codes_z_p"])
            B, T = case["B"], case["T"]
            H.assert_close_frames(npy(zqis)[:, :, ::16, :].reshape(B, -1, T), g["from_codes_z_q_is_sub"].reshape(B, -1, T), what="from_codes z_q_is")
            ozq, _, _ = c_oracle.from_codes(w, g[f"codes_{qi}"])
            if case["Nq"] > 8 or impl == "cuda":  # CUDA-core decode kernel: the oracle's fp32 op order, bit for bit
                assert np.array_equal(npy(zq), ozq), "from_codes (CUDA cores) is bit-exact vs the oracle"
            else:  # tensor-core decode path: same gather, 3xTF32 GEMM accumulation
                H.assert_close_frames(npy(zq), ozq, rtol=5e-6, what="from_codes (tensor cores) vs the oracle")
            lat = torch.from_numpy(g[f"latents_{qi}"]).cuda()
            zq2, zp2, codes2 = m.from_latents(lat)
            assert np.array_equal(npy(codes2), g["from_latents_codes"]), "search on the reference's own z_e must be bit-exact"
            H.assert_close_frames(npy(zq2), g["from_latents_z_q"], what="from_latents z_q")


def test_single_stage_vector_quantize_api():
    case = gi.CASES["cbr_d1024_nq8"]
    m = build_module(case)
    w = H.oracle_weights_for(case)
    z_np = H.latents_for(case)
    z = torch.from_numpy(z_np).cuda()
    z_q, closs, cbloss, idx, z_e = m.quantizers[0](z, loss_per_frame=True)
    o = c_oracle.encode(w, z_np, 1)
    H.assert_codes_match(w, o, npy(idx)[:, None, :])
    assert closs.shape == (case["B"], case["T"]) and z_e.shape == (case["B"], 8, case["T"])
    np.testing.assert_allclose(npy(closs), o["loss_pf"][:, 0, :], rtol=1e-4)
    z_q2, closs2, _, _, _ = m.quantizers[0](z)
    assert closs2.shape == (case["B"],)


def test_mask_utilities_against_reference_fixture():
    import vrvq_b200

    g = H.load_golden("mask_utils")
    x = torch.from_numpy(g["x"]).cuda()
    for nq in (8, 28):
        mk = vrvq_b200.generate_mask_hard(x, nq)
        assert np.array_equal(npy(mk), g[f"mask_nq{nq}"])
        assert vrvq_b200.cal_bpf_from_mask(mk, [10] * nq) == pytest.approx(float(g[f"bpf_nq{nq}"]), rel=1e-6)
        assert vrvq_b200.cal_bpf_from_mask(mk, list(range(1, nq + 1))) == pytest.approx(float(g[f"bpf_ragged_nq{nq}"]), rel=1e-6)
    xi = torch.from_numpy(g["xi"]).cuda()  # the int64 call site, quantize.py:412
    assert np.array_equal(npy(vrvq_b200.generate_mask_hard(xi, 8)), g["mask_int"])
    assert np.array_equal(npy(vrvq_b200.generate_mask_ste(x, 8, alpha=2.0)), g["mask_nq8"])


@pytest.mark.parametrize("T", [1, 3, 31, 32, 33, 64, 87, 88, 431])
def test_edge_frame_counts_against_oracle(T):
    """Ragged tiles, every global-access width (T%4==0 -> 128-bit, T even -> 64-bit, T odd -> 32-bit)."""
    from vrvq_b200 import ops

    case = dict(seed=31, Nq=8, D=1024, K=1024)
    sd = gi.torch_state_dict(gi.make_state_dict(31, 8, 1024))
    w = c_oracle.OracleWeights.from_state_dict(sd)
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    B = 2
    z_np = gi.make_latents(500 + T, B, 1024, T, 0.5)
    imp_np = gi.make_imp_map(600 + T, B, T)
    o = c_oracle.encode(w, z_np, None, imp_np, 0.6)
    out = ops.rvq_encode(pw, torch.from_numpy(z_np).cuda(), None, torch.from_numpy(imp_np).cuda(), 0.6, want_z_q_is=True, want_loss_pf=True)
    excused, skip = H.assert_codes_match(w, o, npy(out.codes))
    assert np.array_equal(npy(out.mask), o["mask"])
    H.assert_close_frames(npy(out.z_q), o["z_q"], skip=skip, what="z_q")
    H.assert_close_frames(npy(out.z_q_is).reshape(B, -1, T), o["z_q_is"].reshape(B, -1, T), skip=skip, what="z_q_is")
    H.assert_close_frames(npy(out.latents), o["latents"], skip=skip, what="latents")
    if excused == 0:
        np.testing.assert_allclose(npy(out.loss_pf), o["loss_pf"], rtol=2e-4, atol=1e-7)
        assert out.loss_sum.item() == pytest.approx(o["loss_masked_sum"], rel=1e-5)
    assert np.array_equal(npy(out.kept), o["kept"])


def test_config2_shape_against_oracle():
    """BASELINE.json configs[1]: B=16 x 10 s (T=862), Nq=8, level sweep {0.25, 0.5, 1.0}; full output dict."""
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(41, 8, 1024))
    w = c_oracle.OracleWeights.from_state_dict(sd)
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    B, T = 16, 862
    z_np = gi.make_latents(42, B, 1024, T, 1.0)
    imp_np = gi.make_imp_map(43, B, T)
    z, imp = torch.from_numpy(z_np).cuda(), torch.from_numpy(imp_np).cuda()
    total_excused = 0
    for level in (0.25, 0.5, 1.0):
        o = c_oracle.encode(w, z_np, None, imp_np, level, want_z_q_is=(level == 1.0))
        out = ops.rvq_encode(pw, z, None, imp, level, want_z_q_is=True)
        excused, skip = H.assert_codes_match(w, o, npy(out.codes))
        total_excused += excused
        assert np.array_equal(npy(out.mask), o["mask"])
        assert np.array_equal(npy(out.kept), o["kept"])
        H.assert_close_frames(npy(out.z_q), o["z_q"], skip=skip, what="z_q")
        if level == 1.0:
            H.assert_close_frames(npy(out.z_q_is).reshape(B, -1, T), o["z_q_is"].reshape(B, -1, T), skip=skip, what="z_q_is")
        # re-mask kernel (scripts/inference.py:95-100) must reproduce the fused z_q from z_q_is
        zq2, mk2, kept2 = ops.remask(out.z_q_is, imp, level * 8)
        ref_mask = c_oracle.generate_mask_hard(imp_np * np.float32(level * 8), 8)
        assert np.array_equal(npy(mk2), ref_mask)
        same = np.array_equal(ref_mask, o["mask"])
        if same:
            # the fused z_q is one tensor-core GEMM over all kept stages, the re-mask sums the stored fp32 stage outputs
            H.assert_close_frames(npy(zq2), npy(out.z_q), what="re-masked z_q vs fused z_q")
            assert np.array_equal(npy(kept2), o["kept"])
    print(f"config-2 shape: {total_excused} audited near-tie frame(s) of {3 * B * T}")


def test_shard_and_view_invariance():
    """Per-frame independence: batch shards, frame-range views and the fused full call agree bit-for-bit."""
    from vrvq_b200 import ops, sharding

    sd = gi.torch_state_dict(gi.make_state_dict(51, 8, 1024))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    B, T = 5, 200
    z = torch.from_numpy(gi.make_latents(52, B, 1024, T, 1.0)).cuda()
    imp = torch.from_numpy(gi.make_imp_map(53, B, T)).cuda()
    lv = torch.tensor([0.2, 0.5, 1.0, 2.0, 0.7], device="cuda")
    full = ops.rvq_encode(pw, z, None, imp, lv, want_z_q_is=True)
    kept = torch.zeros_like(full.kept)
    loss = torch.zeros((1,), dtype=torch.float64, device="cuda")
    for world in (2, 3, 8):
        plan = sharding.plan_shards(B, T, world)
        assert sum(s.frames for segs in plan for s in segs) == B * T
        kept.zero_(); loss.zero_()
        for segs in plan:
            part = sharding.encode_shard(pw, z, segs, 8, imp, lv, want_z_q_is=True)
            for s in segs:
                sl = (slice(s.b, s.b + 1), slice(None), slice(s.t0, s.t1))
                assert torch.equal(part.codes[sl], full.codes[sl])
                assert torch.equal(part.z_q[sl], full.z_q[sl])
                assert torch.equal(part.mask[sl], full.mask[sl])
                assert torch.equal(part.latents[sl], full.latents[sl])
                assert torch.equal(part.z_q_is[s.b, :, :, s.t0:s.t1], full.z_q_is[s.b, :, :, s.t0:s.t1])
            kept += part.kept
            loss += part.loss_sum
        assert torch.equal(kept, full.kept)
        assert loss.item() == pytest.approx(full.loss_sum.item(), rel=1e-9)


def test_error_behaviour():
    import vrvq_b200

    case = gi.CASES["vbr_t1"]
    m = build_module(case)
    z = torch.zeros(1, 1024, 4)
    with pytest.raises(vrvq_b200.VrvqError):  # no CPU fallback
        m(z, level=1.0, imp_map=torch.zeros(1, 1, 4))
    zc = z.cuda()
    with pytest.raises(AssertionError):  # quantize.py:348
        m(zc, n_quantizers=None, level=None, imp_map=torch.zeros(1, 1, 4).cuda())
    with pytest.raises(RuntimeError):  # partial CBR inside the VBR model does not broadcast in the reference
        m(zc, n_quantizers=3)
    r = m(zc, n_quantizers=8)
    assert r["imp_map"] is None and torch.all(r["mask_imp"] == 1)
    m.train()
    with pytest.raises(NotImplementedError):
        m(zc, level=1.0, imp_map=torch.zeros(1, 1, 4).cuda())
    with pytest.raises(NotImplementedError):
        m.from_codes(torch.zeros(1, 8, 4, dtype=torch.int64).cuda())
    cbr = build_module(gi.CASES["cbr_t3"])
    with pytest.raises(IndexError):
        cbr.from_codes(torch.full((1, 8, 4), 1024, dtype=torch.int64).cuda())
    e = cbr(torch.zeros(0, 1024, 7).cuda())
    assert e["codes"].shape == (0, 8, 7) and e["z_q"].shape == (0, 1024, 7)


@pytest.mark.parametrize("D,Nq", [(256, 4), (512, 5), (1024, 32)])
def test_other_input_dims_and_stage_counts_against_oracle(D, Nq):
    """Every kernel instantiation (D in {256, 512, 1024}) and the largest supported stage count, VBR with z_q_is."""
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(100 + D + Nq, Nq, D))
    w = c_oracle.OracleWeights.from_state_dict(sd)
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    B, T = 2, 70
    z_np = gi.make_latents(7 + D, B, D, T, 0.8)
    imp_np = gi.make_imp_map(8 + D, B, T)
    lv = torch.tensor([0.4, 1.3], device="cuda")
    o = c_oracle.encode(w, z_np, None, imp_np, np.array([0.4, 1.3], np.float32))
    out = ops.rvq_encode(pw, torch.from_numpy(z_np).cuda(), None, torch.from_numpy(imp_np).cuda(), lv, want_z_q_is=True)
    excused, skip = H.assert_codes_match(w, o, npy(out.codes))
    assert np.array_equal(npy(out.mask), o["mask"]) and np.array_equal(npy(out.kept), o["kept"])
    H.assert_close_frames(npy(out.z_q), o["z_q"], skip=skip, what="z_q")
    H.assert_close_frames(npy(out.z_q_is).reshape(B, -1, T), o["z_q_is"].reshape(B, -1, T), skip=skip, what="z_q_is")
    H.assert_close_frames(npy(out.latents), o["latents"], skip=skip, what="latents")


def test_unsupported_shapes_fail_loudly():
    import vrvq_b200
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(3, 33, 256))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    with pytest.raises(vrvq_b200.VrvqError, match="n_run"):  # more than 32 stages
        ops.rvq_encode(pw, torch.zeros(1, 256, 8, device="cuda"))
    sd = gi.torch_state_dict(gi.make_state_dict(3, 2, 384))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    assert not pw.supported()
    with pytest.raises(vrvq_b200.VrvqError, match="no kernel"):
        ops.rvq_encode(pw, torch.zeros(1, 384, 8, device="cuda"))
    sd = gi.torch_state_dict(gi.make_state_dict(3, 2, 256))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    with pytest.raises(vrvq_b200.VrvqError):  # non-unit stride along T
        ops.rvq_encode(pw, torch.zeros(1, 256, 16, device="cuda")[:, :, ::2])
    with pytest.raises(vrvq_b200.VrvqError):  # wrong dtype
        ops.rvq_encode(pw, torch.zeros(1, 256, 8, device="cuda", dtype=torch.float16))


def test_padded_z_q_is_rows_extension_gives_identical_values():
    case = gi.CASES["vbr_d1024_nq8"]
    m = build_module(case)
    z = torch.from_numpy(H.latents_for(case)).cuda()
    imp = torch.from_numpy(gi.make_imp_map(case["imp_seed"], case["B"], case["T"])).cuda()
    a = m(z, level=0.5, imp_map=imp)
    m.pad_z_q_is_rows = True
    b = m(z, level=0.5, imp_map=imp)
    assert b["z_q_is"].shape == a["z_q_is"].shape and b["z_q_is"].stride(2) % 32 == 0 and not b["z_q_is"].is_contiguous()
    for k in ("z_q", "z_q_is", "codes", "latents", "mask_imp"):
        assert torch.equal(a[k], b[k]), k


def test_two_devices_in_one_process():
    """the opt-in shared-memory limit is a per-device attribute; a process that drives two GPUs must get it on both
    (the Python API selects the device per call).  Skipped on a one-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    from vrvq_b200 import ops

    sd = gi.torch_state_dict(gi.make_state_dict(61, 8, 1024))
    z_np = gi.make_latents(62, 20, 1024, 300, 1.0)  # large enough for the tensor-core kernel (222 KB of dynamic shared memory)
    imp_np = gi.make_imp_map(63, 20, 300)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        pw = ops.PackedWeights.from_state_dict(sd, dev)
        for impl in ("tc", "cuda"):
            import os
            os.environ["VRVQ_ENCODE_IMPL"] = impl
            try:
                o = ops.rvq_encode(pw, torch.from_numpy(z_np).to(dev), None, torch.from_numpy(imp_np).to(dev), 0.5, want_z_q_is=True)
                zq, _, _ = ops.from_codes(pw, o.codes)
            finally:
                os.environ.pop("VRVQ_ENCODE_IMPL", None)
            torch.cuda.synchronize(dev)
            outs.append((impl, o.codes.cpu(), o.z_q.cpu(), zq.cpu()))
    for a, b in ((outs[0], outs[2]), (outs[1], outs[3])):
        assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]), f"{a[0]}: device 1 differs from device 0"


def test_packed_weight_cache_invalidation():
    """load_state_dict / .to() / in-place optimiser-style updates repack; `.data` surgery needs invalidate_packed()."""
    case = gi.CASES["cbr_t3"]
    m = build_module(case)
    z = torch.from_numpy(H.latents_for(case)).cuda()
    a = m(z)["z_q"].clone()
    with torch.no_grad():
        m.quantizers[3].codebook.weight.mul_(1.5)  # bumps _version of a MIDDLE stage
    b = m(z)["z_q"].clone()
    assert not torch.equal(a, b), "an in-place update of a middle stage must repack"
    m.quantizers[3].codebook.weight.data.mul_(1.0 / 1.5)  # .data: invisible to the version counter ...
    m.invalidate_packed()  # ... hence the explicit hook
    c = m(z)["z_q"]
    assert torch.allclose(a, c, rtol=1e-4, atol=1e-5) and not torch.equal(b, c)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd["quantizers.5.out_proj.bias"] += 1.0
    m.load_state_dict(sd)
    d = m(z)["z_q"]
    assert not torch.equal(c, d), "load_state_dict must repack"


def test_level_tensor_shapes_follow_the_reference_broadcast():
    """quantize.py:389: `imp_map [B,1,T] * level`: [B,1,1] is per item, [1,1,T] per frame; a shape-[B] level with B != T is a
    broadcast error in the reference and must not be silently taken as per-item levels."""
    from vrvq_b200 import ops

    case = gi.CASES["vbr_tensor_level"]  # B=3, T=21
    m = build_module(case)
    B, T, Nq = case["B"], case["T"], case["Nq"]
    z = torch.from_numpy(H.latents_for(case)).cuda()
    imp = torch.from_numpy(gi.make_imp_map(case["imp_seed"], B, T)).cuda()
    per_frame = torch.linspace(0.2, 2.0, T, device="cuda").view(1, 1, T)
    r = m(z, level=per_frame, imp_map=imp)
    assert torch.equal(r["mask_imp"], ops.generate_mask_hard(imp * per_frame * Nq, Nq))
    per_item = torch.tensor([0.3, 1.0, 2.5], device="cuda")
    r3 = m(z, level=per_item.view(B, 1, 1), imp_map=imp)
    assert torch.equal(r3["mask_imp"], ops.generate_mask_hard(imp * per_item.view(B, 1, 1) * Nq, Nq))
    with pytest.raises(RuntimeError):
        m(z, level=per_item, imp_map=imp)  # [3] against [3,1,21]: the reference's broadcast fails
    sq = torch.linspace(0.5, 1.5, T, device="cuda")  # shape [T]: broadcasts along frames, like the reference
    r1 = m(z, level=sq, imp_map=imp)
    assert torch.equal(r1["mask_imp"], ops.generate_mask_hard(imp * sq * Nq, Nq))
