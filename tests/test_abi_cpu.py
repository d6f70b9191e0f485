"""CPU: the C-ABI library loads, exports every symbol include/vrvq.h declares, packs weights with the
reference's arithmetic, validates arguments, and fails loudly without a device (no compute happens here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests.golden import gen_inputs as gi
from vrvq_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "vrvq.h")).read()
    declared = set(re.findall(r"\b(vrvq_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), f"binding list out of sync with the header: {declared ^ set(_lib.EXPORTS)}"
    for name in declared:
        assert hasattr(L, name), f"libvrvq.so does not export {name}"
    assert L.vrvq_abi_version() == 1


def test_struct_layouts_match_the_header():
    # the library rejects a struct_size it was not compiled with, which catches binding/ABI drift
    L = _lib.lib()
    a = _lib.EncodeArgs()
    a.struct_size = C.sizeof(_lib.EncodeArgs) - 8
    assert L.vrvq_rvq_encode_f32(C.byref(a), None) == -1
    assert b"struct_size" in L.vrvq_last_error()
    f = _lib.FromCodesArgs()
    f.struct_size = 4
    # device check happens first for from_codes on a GPU-less box; either error is loud
    assert L.vrvq_from_codes_f32(C.byref(f), None) in (-1, -4)


def test_pack_weights_reproduces_torch_normalize():
    sd = gi.torch_state_dict(gi.make_state_dict(5, 3, 512))
    pw = ops.PackedWeights.from_state_dict(sd, "cpu")
    ow = c_oracle.OracleWeights.from_state_dict(sd)
    for s in range(3):
        cb, c2 = pw.normalized_codebook(s)
        ref = torch.nn.functional.normalize(sd[f"quantizers.{s}.codebook.weight"])
        assert np.array_equal(cb, ref.numpy()), "F.normalize(codebook) must be bit-identical (quantize.py:93)"
        assert np.array_equal(c2, ref.pow(2).sum(1).numpy()), "codebook.pow(2).sum(1) must be bit-identical (quantize.py:99)"
        assert np.array_equal(cb, ow.cb_nrm[s]) and np.array_equal(c2, ow.c2[s])
    assert pw.supported()
    assert _lib.lib().vrvq_supported(1024, 1024, 8) == 1 and _lib.lib().vrvq_supported(1000, 1024, 8) == 0
    assert _lib.lib().vrvq_supported(1024, 1024, 16) == 0


def test_fold_matches_reference_weight_norm():
    v, g = torch.randn(8, 1024, 1), torch.rand(8, 1, 1) + 0.5
    conv = torch.nn.utils.weight_norm(torch.nn.Conv1d(1024, 8, 1))
    with torch.no_grad():
        conv.weight_v.copy_(v)
        conv.weight_g.copy_(g)
        conv(torch.zeros(1, 1024, 1))  # runs the pre-forward hook
    assert torch.equal(conv.weight, ops.fold_weight_norm(v, g))


def test_argument_validation():
    L = _lib.lib()
    assert L.vrvq_blob_bytes(8, 1024, 1024, 4) == 0
    assert L.vrvq_pack_weights(1, 1024, 1024, 8, None, None, None, None, None, None, 0) == -1
    assert L.vrvq_pack_weights(1, 1024, 1024, 7, None, None, None, None, None, None, 0) == -2
    a = _lib.EncodeArgs()
    a.struct_size = C.sizeof(_lib.EncodeArgs)
    a.B, a.T, a.input_dim, a.n_codebooks, a.codebook_size, a.n_run = 1, 4, 1024, 8, 1024, 9
    assert L.vrvq_rvq_encode_f32(C.byref(a), None) == -1 and b"n_run" in L.vrvq_last_error()
    a.n_run = 8
    a.input_dim = 1000
    assert L.vrvq_rvq_encode_f32(C.byref(a), None) == -2
    a.input_dim = 1024
    assert L.vrvq_rvq_encode_f32(C.byref(a), None) == -1  # NULL blob / z / codes
    assert L.vrvq_generate_mask_hard_f32(None, 0, 1, 4, 8, None, 0, 0, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    import vrvq_b200

    m = vrvq_b200.ResidualVectorQuantize(input_dim=1024, n_codebooks=2).eval()
    with pytest.raises(vrvq_b200.VrvqError):
        m(torch.zeros(1, 1024, 4))
    with pytest.raises(vrvq_b200.VrvqError):
        vrvq_b200.generate_mask_hard(torch.zeros(1, 1, 4), 8)
    with pytest.raises(vrvq_b200.VrvqError):
        vrvq_b200.cal_bpf_from_mask(torch.zeros(1, 8, 4), [10] * 8)
    # the library itself refuses compute without an sm_100 device
    L = _lib.lib()
    x = np.zeros((1, 4), np.float32)
    m_ = np.zeros((1, 8, 4), np.float32)
    assert L.vrvq_generate_mask_hard_f32(x.ctypes.data, 4, 1, 4, 8, m_.ctypes.data, 32, 4, None) == -4
    assert b"no CPU fallback" in L.vrvq_last_error() or b"sm_100a" in L.vrvq_last_error()


def test_module_surface_matches_reference_signatures():
    import inspect

    import vrvq_b200

    sig = inspect.signature(vrvq_b200.VBRResidualVectorQuantize.forward)
    assert list(sig.parameters)[:5] == ["self", "z", "n_quantizers", "feat_enc", "level"]
    sig = inspect.signature(vrvq_b200.ResidualVectorQuantize.forward)
    assert list(sig.parameters) == ["self", "z", "n_quantizers"]
    sig = inspect.signature(vrvq_b200.DAC_VRVQ.encode)
    assert list(sig.parameters) == ["self", "audio_data", "n_quantizers", "level"] and sig.parameters["level"].default == 1
    m = vrvq_b200.ResidualVectorQuantize()  # reference defaults quantize.py:112-119
    assert (m.input_dim, m.n_codebooks, m.codebook_size) == (512, 9, 1024)
    keys = list(m.state_dict().keys())[:7]
    assert keys == [f"quantizers.0.{k}" for k in ("in_proj.bias", "in_proj.weight_g", "in_proj.weight_v", "out_proj.bias",
                                                     "out_proj.weight_g", "out_proj.weight_v", "codebook.weight")]
    with pytest.raises(vrvq_b200.VrvqError):
        vrvq_b200.ResidualVectorQuantize(codebook_dim=16)
    m.train()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 512, 4))
