"""Container-only (needs the reference checkout): oracle vs the live reference on larger random cases than the
committed fixtures, with the near-tie audit, plus state-dict compatibility of the Python mirror."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, ref_import
from tests import helpers as H
from tests.golden import gen_inputs as gi

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")]


@pytest.fixture(scope="module")
def ref():
    return ref_import.load()


def test_cbr_random_init_large(ref):
    """Reference-initialised weights (torch.manual_seed), B=4 x T=431, Nq=8: 13.8k frame-stages."""
    torch.manual_seed(0)
    m = ref.ResidualVectorQuantize(input_dim=1024, n_codebooks=8, codebook_size=1024, codebook_dim=8).eval()
    z = torch.randn(4, 1024, 431)
    with torch.no_grad():
        r = m(z)
    w = c_oracle.OracleWeights.from_state_dict(m.state_dict())
    o = c_oracle.encode(w, z.numpy(), None, want_z_q_is=False)
    excused, mask = H.assert_codes_match(w, o, r["codes"].numpy())
    H.assert_close_frames(o["z_q"], r["z_q"].numpy(), skip=mask, what="z_q")
    H.assert_close_frames(o["latents"], r["latents"].numpy(), skip=mask, what="latents", rtol=1e-4)
    assert o["commitment_loss"] == pytest.approx(r["commitment_loss"].item(), rel=1e-4)


def test_vbr_with_real_subnet(ref):
    torch.manual_seed(1)
    m = ref.VBRResidualVectorQuantize(input_dim=1024, n_codebooks=8, codebook_size=1024, codebook_dim=8, level_min=0.125,
                                      level_max=6.0, imp2mask_alpha=2.0).eval()
    z, feat = torch.randn(2, 1024, 87), torch.randn(2, 1024, 87)
    for level in (0.25, 1.0, 3.0):
        with torch.no_grad():
            r = m(z, n_quantizers=None, feat_enc=feat, level=level)
        w = c_oracle.OracleWeights.from_state_dict(m.state_dict())
        o = c_oracle.encode(w, z.numpy(), None, r["imp_map"].numpy(), level, want_z_q_is=False)
        H.assert_codes_match(w, o, r["codes"].numpy())
        assert np.array_equal(o["mask"], r["mask_imp"].numpy())
        H.assert_close_frames(o["z_q"], r["z_q"].numpy(), what="z_q")
        assert ref.cal_bpf_from_mask(r["mask_imp"], [10] * 8) == pytest.approx(float((o["kept"] * 10).sum()) / (2 * 87), rel=1e-6)


def test_mirror_state_dict_keys_match_reference(ref):
    """The Python mirror must load a reference checkpoint unchanged (checkpoint row)."""
    import vrvq_b200

    torch.manual_seed(2)
    r = ref.VBRResidualVectorQuantize(input_dim=1024, n_codebooks=8, codebook_size=1024, codebook_dim=8, level_min=0.125, level_max=6.0)
    ours = vrvq_b200.VBRResidualVectorQuantize(input_dim=1024, n_codebooks=8, codebook_size=1024, codebook_dim=8, level_min=0.125, level_max=6.0)
    rsd, osd = r.state_dict(), ours.state_dict()
    assert list(rsd.keys()) == list(osd.keys())
    assert all(rsd[k].shape == osd[k].shape for k in rsd)
    ours.load_state_dict(rsd, strict=True)
    # the PyTorch importance subnet (upstream producer) must agree numerically with the reference's
    feat = torch.randn(2, 1024, 50)
    with torch.no_grad():
        a, b = r.imp_subnet(feat), ours.imp_subnet.forward_torch(feat)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


def test_mirror_encoder_matches_reference(ref):
    import vrvq_b200

    torch.manual_seed(3)
    r = ref.DAC_VRVQ(n_codebooks=8, model_type="VBR", level_min=0.125, level_max=6.0).eval()
    ours = vrvq_b200.DAC_VRVQ(n_codebooks=8, model_type="VBR", level_min=0.125, level_max=6.0).eval()
    missing, unexpected = [], []
    sd = {k: v for k, v in r.state_dict().items() if not k.startswith("decoder.")}
    assert list(sd.keys()) == list(ours.state_dict().keys())
    ours.load_reference_state_dict(r.state_dict())
    x = torch.randn(1, 1, 44100) * 0.1
    with torch.no_grad():
        xr = r.preprocess(x, 44100)
        xo = ours.preprocess(x, 44100)
        assert torch.equal(xr, xo) and xo.shape[-1] == 44544
        zr, fr = r.encoder(xr, return_feat=True)
        zo, fo = ours.encoder(xo, return_feat=True)
    assert zr.shape == (1, 1024, 87)
    assert torch.allclose(zr, zo, rtol=1e-4, atol=1e-6) and torch.allclose(fr, fo, rtol=1e-4, atol=1e-5)


def test_dac_file_interoperates_with_the_reference_container(ref, tmp_path):
    """models/dac_base.py:18-58: files written by the reference's DACFile load through the mirror and vice versa, and the
    oracle's packed codes are the uint16 array the reference stores."""
    from oracle import wire as ow
    from vrvq_b200 import wire

    codes = torch.randint(0, 1024, (2, 8, 60))
    meta = dict(chunk_length=60, original_length=30720, input_db=torch.tensor([-18.5]), channels=1, sample_rate=44100, padding=True,
                dac_version="1.0.0")
    p_ref = ref.DACFile(codes=codes, **meta).save(tmp_path / "from_ref")
    stored = np.load(p_ref, allow_pickle=True)[()]["codes"]
    assert stored.dtype == np.uint16 and np.array_equal(stored, ow.pack_codes(codes.numpy())[0])
    ours = wire.DACFile.load(p_ref)
    assert torch.equal(ours.codes, codes) and ours.counts is None and ours.original_length == 30720 and ours.padding is True
    p_ours = wire.DACFile(codes=codes, **meta).save(tmp_path / "from_ours")
    theirs = ref.DACFile.load(p_ours)
    assert torch.equal(theirs.codes, codes) and theirs.chunk_length == 60 and theirs.sample_rate == 44100
    assert open(p_ref, "rb").read() == open(p_ours, "rb").read(), "CBR files are byte-identical to the reference's"


class _FixedImportance(torch.nn.Module):
    """Stands in for the importance subnet (an upstream producer of the path) so that every mask edge is exercised."""

    def __init__(self, imp):
        super().__init__()
        self.imp = imp

    def forward(self, feat):
        return self.imp


@pytest.mark.parametrize("B", [4, 16])
def test_torch_port_is_bit_identical_to_the_live_reference(ref, B):
    """oracle/torch_port.py is what bench.py's `cpu_baseline` leg and `--impl reference` arm time when the reference checkout
    is absent (the GPU box): it must be the reference's computation, bit for bit, at the bench's own shapes
    (B in {4, 16} x T = 862, D = 1024, Nq = 8; VBR with the level sweep of configs[1], and CBR with early exit)."""
    from oracle import torch_port

    torch.set_num_threads(8)
    T, D, Nq = 862, 1024, 8
    sd = gi.torch_state_dict(gi.make_state_dict(0, Nq, D, 1024))  # the bench's weights (bench.py: make_state)
    z = torch.randn(B, D, T, generator=torch.Generator().manual_seed(1234))
    imp = torch.rand(B, 1, T, generator=torch.Generator().manual_seed(4321))
    w = torch_port.TorchPortWeights(sd)
    vbr = ref.VBRResidualVectorQuantize(input_dim=D, n_codebooks=Nq, codebook_size=1024, codebook_dim=8, level_min=0.125, level_max=6.0,
                                        imp2mask_alpha=2.0).eval()
    missing, unexpected = vbr.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("imp_subnet.") for k in missing)
    vbr.imp_subnet = _FixedImportance(imp)
    for level in (0.25, 1.0):
        with torch.no_grad():
            r = vbr(z, n_quantizers=None, feat_enc=z, level=level)
        p = torch_port.rvq_forward(w, z, None, imp, level)
        for k in ("codes", "z_q", "z_q_is", "latents", "mask_imp", "commitment_loss", "codebook_loss"):
            assert torch.equal(p[k], r[k]), f"VBR level {level}: {k} differs from the reference"
    cbr = ref.ResidualVectorQuantize(input_dim=D, n_codebooks=Nq, codebook_size=1024, codebook_dim=8).eval()
    cbr.load_state_dict(sd, strict=True)
    for nq in (None, 3):
        with torch.no_grad():
            r = cbr(z, n_quantizers=nq)
        p = torch_port.rvq_forward(w, z, nq)
        for k in ("codes", "z_q", "latents", "commitment_loss", "codebook_loss"):
            assert torch.equal(p[k], r[k]), f"CBR n_quantizers={nq}: {k} differs from the reference"
