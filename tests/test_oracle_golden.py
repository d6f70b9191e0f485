"""CPU: the C restatement (oracle/rvq_oracle.c) against fixtures produced by the real reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import c_oracle
from tests import helpers as H
from tests.golden import gen_inputs as gi

VBR = [n for n, c in gi.CASES.items() if c["kind"] == "vbr"]
CBR = [n for n, c in gi.CASES.items() if c["kind"] == "cbr"]


@pytest.mark.parametrize("name", VBR)
def test_vbr_forward_matches_reference(name):
    case, g = gi.CASES[name], H.load_golden(name)
    w, z = H.oracle_weights_for(case), H.latents_for(case)
    imp = gi.make_imp_map(case["imp_seed"], case["B"], case["T"])
    for li, level in enumerate(case["levels"]):
        o = c_oracle.encode(w, z, None, imp, level)
        assert np.array_equal(o["codes"], g["codes"]), "codes must be bit-identical to the reference"
        assert np.array_equal(o["mask"], g[f"mask_{li}"]), "hard mask must be bit-identical"
        H.assert_close_frames(o["z_q"], g[f"z_q_{li}"], what="z_q")
        H.assert_close_frames(o["latents"], g["latents"], what="latents")
        B, T = case["B"], case["T"]
        H.assert_close_frames(o["z_q_is"][:, :, ::16, :].reshape(B, -1, T), g["z_q_is_sub"].reshape(B, -1, T), what="z_q_is")
        assert o["commitment_loss"] == pytest.approx(float(g[f"commitment_loss_{li}"]), rel=1e-5)
        assert o["codebook_loss"] == pytest.approx(float(g[f"codebook_loss_{li}"]), rel=1e-5)
        bpf = float((o["kept"] * 10).sum()) / (B * T)
        assert bpf == pytest.approx(float(g[f"bpf_{li}"]), rel=1e-6)
        assert c_oracle.cal_bpf_from_mask(o["mask"], [10] * case["Nq"]) == pytest.approx(bpf, rel=1e-12)


@pytest.mark.parametrize("name", CBR)
def test_cbr_forward_and_decode_match_reference(name):
    case, g = gi.CASES[name], H.load_golden(name)
    w, z = H.oracle_weights_for(case), H.latents_for(case)
    for qi, nq in enumerate(case["n_quantizers"]):
        o = c_oracle.encode(w, z, nq)
        assert np.array_equal(o["codes"], g[f"codes_{qi}"])
        assert np.all(o["mask"] == 1.0)
        H.assert_close_frames(o["z_q"], g[f"z_q_{qi}"], what="z_q")
        H.assert_close_frames(o["latents"], g[f"latents_{qi}"], what="latents")
        assert o["commitment_loss"] == pytest.approx(float(g[f"commitment_loss_{qi}"]), rel=1e-5)
        if nq is None:
            zq, zp, zqis = c_oracle.from_codes(w, g[f"codes_{qi}"], True)
            H.assert_close_frames(zq, g["from_codes_z_q"], what="from_codes z_q")
            assert np.array_equal(zp, g["from_codes_z_p"])
            B, T = case["B"], case["T"]
            H.assert_close_frames(zqis[:, :, ::16, :].reshape(B, -1, T), g["from_codes_z_q_is_sub"].reshape(B, -1, T), what="from_codes z_q_is")
            zq2, zp2, codes2 = c_oracle.from_latents(w, g[f"latents_{qi}"])
            assert np.array_equal(codes2, g["from_latents_codes"]), "decode_latents on the reference's own z_e must be bit-exact"
            H.assert_close_frames(zq2, g["from_latents_z_q"], what="from_latents z_q")


def test_mask_utilities_match_reference():
    g = H.load_golden("mask_utils")
    for nq in (8, 28):
        m = c_oracle.generate_mask_hard(g["x"], nq)
        assert np.array_equal(m, g[f"mask_nq{nq}"])
        assert c_oracle.cal_bpf_from_mask(m, [10] * nq) == pytest.approx(float(g[f"bpf_nq{nq}"]), rel=1e-6)
        assert c_oracle.cal_bpf_from_mask(m, list(range(1, nq + 1))) == pytest.approx(float(g[f"bpf_ragged_nq{nq}"]), rel=1e-6)
    assert np.array_equal(c_oracle.generate_mask_hard(g["xi"].astype(np.float32), 8), g["mask_int"])


def test_empty_and_degenerate_inputs():
    case = gi.CASES["cbr_t3"]
    w = H.oracle_weights_for(case)
    o = c_oracle.encode(w, np.zeros((0, 1024, 5), np.float32), None)
    assert o["codes"].shape == (0, 8, 5)
    o = c_oracle.encode(w, np.zeros((2, 1024, 0), np.float32), None)
    assert o["z_q"].shape == (2, 1024, 0)
    # all-zero latent: z_e = b_in, everything finite, deterministic
    o = c_oracle.encode(w, np.zeros((1, 1024, 2), np.float32), None)
    assert np.isfinite(o["z_q"]).all() and np.array_equal(o["codes"][0, :, 0], o["codes"][0, :, 1])
