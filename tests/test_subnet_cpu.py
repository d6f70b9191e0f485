"""Importance subnet without a GPU: the numpy oracle against the fixtures generated from the
reference (tests/golden/subnet_*.npz, make_golden.py) and against the live reference when its checkout is present; the host
side of the C ABI (weight packing layout); the module mirror's state-dict layout and error behaviour."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ref_import, subnet_port as sp
from tests import helpers as H
from tests.golden import gen_inputs as gi

IMP_ATOL = 5e-6  # |imp_map - reference|: fp32 with an unspecified (oneDNN) reduction order over up to 3072 terms per layer


def case_inputs(c):
    sd = gi.make_subnet_state_dict(c["seed"], c["d_input"], c["d_feat"], c["widths"])
    x = gi.make_latents(c["seed"] + 1000, c["B"], c["d_input"], c["T"], c["sigma"])
    return sd, x


@pytest.mark.parametrize("name", list(gi.SUBNET_CASES))
def test_oracle_matches_reference_fixture(name):
    c = gi.SUBNET_CASES[name]
    sd, x = case_inputs(c)
    g = H.load_golden(name)["imp_map"]
    assert g.shape == (c["B"], 1, c["T"]) and g.dtype == np.float32
    for dt in (np.float32, np.float64):
        o = sp.importance_subnet(sd, x, dtype=dt)
        assert o.shape == g.shape
        assert np.abs(o.astype(np.float64) - g).max() <= IMP_ATOL, (name, dt)


@pytest.mark.reference
@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")
def test_oracle_matches_live_reference_random_init():
    """The reference's own initialisation (alpha = 1, g = ||v||, kaiming v) on a longer sequence than the fixtures."""
    ref_import.load()
    from models.importance_subnet import ImportanceSubnet

    torch.manual_seed(5)
    m = ImportanceSubnet(d_input=1024, d_feat=1024).eval()
    x = torch.randn(2, 1024, 300)
    with torch.no_grad():
        r = m(x).numpy()
    sd = {k: v.numpy() for k, v in m.state_dict().items()}
    assert np.abs(sp.importance_subnet(sd, x.numpy(), dtype=np.float32) - r).max() <= IMP_ATOL


def test_pack_conv3_layout():
    """vrvq_pack_conv3_weights (host-only entry point): packed[(ci*3+k) * Cout_padded + co] = w[co,ci,k], zero filled."""
    from vrvq_b200 import _lib

    L = _lib.lib()
    rng = np.random.Generator(np.random.PCG64(3))
    for cout, cin in [(1, 8), (130, 16), (128, 24)]:
        w = rng.normal(size=(cout, cin, 3)).astype(np.float32)
        n = L.vrvq_conv3_packed_floats(cout, cin)
        cp = -(-cout // 128) * 128
        assert n == cin * 3 * cp
        out = np.full(n, 7.0, np.float32)
        assert L.vrvq_pack_conv3_weights(cout, cin, w.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n) == 0
        p = out.reshape(cin, 3, cp)
        assert np.array_equal(p[:, :, :cout], w.transpose(1, 2, 0)) and not p[:, :, cout:].any()
        assert L.vrvq_pack_conv3_weights(cout, cin, w.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n - 1) == -1
    assert L.vrvq_conv3_packed_floats(0, 8) == 0


def test_mirror_layout_and_errors():
    from vrvq_b200 import VrvqError
    from vrvq_b200.layers import ImportanceSubnet

    c = gi.SUBNET_CASES["subnet_small_t3"]
    sd, x = case_inputs(c)
    m = ImportanceSubnet(d_input=c["d_input"], d_feat=c["d_feat"], intermediate_channels=list(c["widths"]))
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(gi.torch_state_dict(sd), strict=True)
    # the differentiable PyTorch formulation (training side) agrees with the reference fixture
    with torch.no_grad():
        y = m.forward_torch(torch.from_numpy(x)).numpy()
    assert np.abs(y - H.load_golden("subnet_small_t3")["imp_map"]).max() <= IMP_ATOL
    # eval forward is the CUDA path only: no CPU fallback
    with pytest.raises(VrvqError):
        m.eval()(torch.from_numpy(x))
