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


def _tc_layout(Nq, D):
    """Mirror of csrc/common.cuh TcLayout (float offsets inside the TC section)."""
    nch, nj, ngrp = D // 32, D // 128, (Nq + 7) // 8
    vp = lambda g: (2 * g + 3) // 4 * 4
    gx_base = lambda g: sum(vp(h) for h in range(g))
    L = dict(nch=nch, nj=nj, ngrp=ngrp, vp=vp, gx_base=gx_base, off_win=0)
    L["off_gx"] = ngrp * nch * 4096
    L["off_wout"] = L["off_gx"] + gx_base(ngrp) * 4096
    L["off_bout"] = L["off_wout"] + Nq * nj * 3072
    L["off_gg"] = L["off_bout"] + ngrp * nj * 2048
    L["off_bin"] = L["off_gg"] + (Nq * (Nq - 1) // 2 * 72 + 3) // 4 * 4
    L["off_spc"] = L["off_bin"] + 2 * Nq * 8
    L["off_cbk"] = L["off_spc"] + Nq * 16
    L["total"] = L["off_cbk"] + Nq * 9216
    return L


def _tile_get(tile, rows, r, k):  # canonical K-major, no swizzle: 8-row x 16-byte core matrices
    return tile[((k // 4) * rows + r) * 4 + (k % 4)]


def _folded(sd, Nq):
    w_in = np.stack([ops.fold_weight_norm(sd[f"quantizers.{s}.in_proj.weight_v"], sd[f"quantizers.{s}.in_proj.weight_g"])[:, :, 0].numpy() for s in range(Nq)])
    w_out = np.stack([ops.fold_weight_norm(sd[f"quantizers.{s}.out_proj.weight_v"], sd[f"quantizers.{s}.out_proj.weight_g"])[:, :, 0].numpy() for s in range(Nq)])
    b_in = np.stack([sd[f"quantizers.{s}.in_proj.bias"].numpy() for s in range(Nq)])
    b_out = np.stack([sd[f"quantizers.{s}.out_proj.bias"].numpy() for s in range(Nq)])
    return w_in, b_in, w_out, b_out


def test_tensor_core_blob_section_layout():
    """The tensor-core section of the packed blob (csrc/common.cuh TcLayout): TF32 head/remainder split is exact, tiles are in
    the canonical K-major UMMA layout, the 8x8 correction matrices equal W_in[s] @ W_out[j] in binary64."""
    Nq, D, K = 3, 256, 1024
    sd = gi.torch_state_dict(gi.make_state_dict(11, Nq, D))
    pw = ops.PackedWeights.from_state_dict(sd, "cpu")
    blob = pw.host_blob
    hdr = blob[:16].view(np.int32)
    tc_off, tc_floats = int(hdr[6]), int(hdr[7])
    assert tc_off > 0 and tc_off + tc_floats == blob.size, "TC section is appended after the per-stage sections"
    tc = blob[tc_off:]
    w_in, b_in, w_out, b_out = _folded(sd, Nq)
    L = _tc_layout(Nq, D)
    nj = L["nj"]
    off_win, off_wout, off_gg, off_bin, off_cbk = L["off_win"], L["off_wout"], L["off_gg"], L["off_bin"], L["off_cbk"]
    assert L["total"] == tc_floats and L["off_gx"] == L["off_wout"], "one stage group: no virtual chunks"
    tile_get = _tile_get

    # WIN chunk 1: rows 0..63 heads, 64..127 remainders of W_in[s][oc][32 + k]
    tile = tc[off_win + 4096: off_win + 2 * 4096]
    for s, oc, k in [(0, 0, 0), (1, 3, 17), (2, 7, 31)]:
        x = w_in[s, oc, 32 + k]
        h, l = tile_get(tile, 128, 8 * s + oc, k), tile_get(tile, 128, 64 + 8 * s + oc, k)
        assert (np.float32(h).view(np.uint32) & 0x1FFF) == 0, "head must be a TF32 value (13 low mantissa bits zero)"
        assert np.float32(h) + np.float32(l) == x and abs(l) <= abs(x) * 2.0 ** -11
    assert tile_get(tile, 128, 8 * Nq, 0) == 0.0, "rows above 8*Nq are zero"
    # WOUT chunk (s=1, j=1): row i <-> channel 128 j + 4 (i % 32) + i // 32; bias tile k=0 head, k=1 remainder
    chunk = tc[off_wout + (1 * nj + 1) * 3072: off_wout + (1 * nj + 2) * 3072]
    for i, k in [(0, 0), (33, 5), (127, 7)]:
        ch = 128 + 4 * (i % 32) + i // 32
        assert np.float32(tile_get(chunk[:1024], 128, i, k)) + np.float32(tile_get(chunk[1024:2048], 128, i, k)) == w_out[1, ch, k]
        assert np.float32(tile_get(chunk[2048:], 128, i, 0)) + np.float32(tile_get(chunk[2048:], 128, i, 1)) == b_out[1, ch]
        assert tile_get(chunk[2048:], 128, i, 2) == 0.0
    # GG pair (j=0, s=2): G = W_in[2] @ W_out[0], g = W_in[2] @ b_out[0]
    pair = 0 * Nq - 0 + (2 - 0 - 1)
    G = tc[off_gg + pair * 72: off_gg + pair * 72 + 72]
    ref = w_in[2].astype(np.float64) @ w_out[0].astype(np.float64)
    np.testing.assert_array_equal(G[:64].reshape(8, 8), ref.astype(np.float32))
    np.testing.assert_array_equal(G[64:], (w_in[2].astype(np.float64) @ b_out[0].astype(np.float64)).astype(np.float32))
    # BIN: b_in, then b_in' (identical for a single stage group)
    np.testing.assert_array_equal(tc[off_bin: off_bin + Nq * 8].reshape(Nq, 8), b_in)
    np.testing.assert_array_equal(tc[off_bin + Nq * 8: off_bin + 2 * Nq * 8].reshape(Nq, 8), b_in)
    # CBK stage 2: the normalised codebook as a [2][K][4] tile + c2
    cb, c2 = pw.normalized_codebook(2)
    cbk = tc[off_cbk + 2 * 9216: off_cbk + 3 * 9216]
    assert np.array_equal(cbk[:4096].reshape(K, 4), cb[:, :4]) and np.array_equal(cbk[4096:8192].reshape(K, 4), cb[:, 4:])
    assert np.array_equal(cbk[8192:], c2)


def test_tensor_core_blob_section_stage_groups():
    """Models with more than 8 codebooks (conf/base_24kbps.yml: 28) run in groups of 8 stages: per-group W_in rows, the
    cross-group corrections as "virtual channel" chunks GX (B = -W_in[s] @ W_out[j] for the stages j of earlier groups, zero
    elsewhere), their bias terms folded into b_in', one bias tile per group."""
    Nq, D = 19, 256
    sd = gi.torch_state_dict(gi.make_state_dict(13, Nq, D))
    pw = ops.PackedWeights.from_state_dict(sd, "cpu")
    hdr = pw.host_blob[:16].view(np.int32)
    tc = pw.host_blob[int(hdr[6]):]
    L = _tc_layout(Nq, D)
    assert L["total"] == int(hdr[7]) and L["ngrp"] == 3 and L["gx_base"](3) == 8
    w_in, b_in, w_out, b_out = _folded(sd, Nq)
    G = lambda s, j: w_in[s].astype(np.float64) @ w_out[j].astype(np.float64)  # [c][k]
    # WIN group 2, chunk 3: local rows 8 (s - 16) + oc; rows of the absent stages 19..23 are zero
    tile = tc[(2 * L["nch"] + 3) * 4096: (2 * L["nch"] + 4) * 4096]
    x = w_in[18, 5, 3 * 32 + 9]
    assert np.float32(_tile_get(tile, 128, 8 * 2 + 5, 9)) + np.float32(_tile_get(tile, 128, 64 + 8 * 2 + 5, 9)) == x
    assert _tile_get(tile, 128, 8 * 3, 0) == 0.0 and _tile_get(tile, 128, 64 + 8 * 7 + 7, 31) == 0.0
    # GX group 1 (stages 8..15): 4 chunks, chunk v covers the earlier stages j = 4v .. 4v+3 (< 8): chunks 2, 3 are padding
    for v, jj, s, c, kk in [(0, 0, 8, 0, 0), (1, 3, 15, 7, 7), (0, 2, 12, 3, 5)]:
        tile = tc[L["off_gx"] + (L["gx_base"](1) + v) * 4096:][:4096]
        want = np.float32(-G(s, 4 * v + jj)[c, kk])
        got = np.float32(_tile_get(tile, 128, 8 * (s - 8) + c, 8 * jj + kk)) + np.float32(_tile_get(tile, 128, 64 + 8 * (s - 8) + c, 8 * jj + kk))
        assert got == want
    assert not tc[L["off_gx"] + (L["gx_base"](1) + 2) * 4096:][:2 * 4096].any(), "padding chunks are zero"
    # GX group 2 (stages 16..18): chunk 3 covers stages 12..15; rows of stages >= 19 zero
    tile = tc[L["off_gx"] + (L["gx_base"](2) + 3) * 4096:][:4096]
    want = np.float32(-G(17, 14)[2, 6])
    assert np.float32(_tile_get(tile, 128, 8 * 1 + 2, 8 * 2 + 6)) + np.float32(_tile_get(tile, 128, 64 + 8 * 1 + 2, 8 * 2 + 6)) == want
    assert _tile_get(tile, 128, 8 * 3 + 0, 0) == 0.0
    # b_in'[s] = b_in[s] - sum_{j < 8 (s // 8)} W_in[s] @ b_out[j]
    bp = tc[L["off_bin"] + Nq * 8: L["off_bin"] + 2 * Nq * 8].reshape(Nq, 8)
    np.testing.assert_array_equal(bp[:8], b_in[:8])
    for s in (8, 18):
        want = b_in[s].astype(np.float64) - sum(w_in[s].astype(np.float64) @ b_out[j].astype(np.float64) for j in range(8 * (s // 8)))
        np.testing.assert_array_equal(bp[s], want.astype(np.float32))
    # BOUT group 2, chunk 1: element (row i, k) = b_out[16 + k][channel(i)] (head + remainder), k >= 3 zero
    bt = tc[L["off_bout"] + (2 * L["nj"] + 1) * 2048:][:2048]
    i, k = 37, 2
    ch = 128 + 4 * (i % 32) + i // 32
    assert np.float32(_tile_get(bt[:1024], 128, i, k)) + np.float32(_tile_get(bt[1024:], 128, i, k)) == b_out[16 + k, ch]
    assert _tile_get(bt[:1024], 128, i, 3) == 0.0


def test_blob_without_tensor_core_section_for_large_nq():
    sd = gi.torch_state_dict(gi.make_state_dict(12, 33, 256))
    pw = ops.PackedWeights.from_state_dict(sd, "cpu")
    hdr = pw.host_blob[:16].view(np.int32)
    assert int(hdr[6]) == 0 and int(hdr[7]) == 0, "Nq > 32 has no tensor-core kernel: no TC section"


def test_flat_tile_order_is_a_permutation_with_the_spanning_tiles_first():
    """csrc/common.cuh:flat_tile_at (the order in which the flat-tiling kernel's CTAs walk their tiles) through its diagnostic entry:
    a permutation of [0, ceil(B T / 128)) whose first B - 1 positions are the tiles that hold an item boundary."""
    L = _lib.lib()
    L.vrvq_flat_tile_order.argtypes = [C.c_int, C.c_int, C.c_int]
    L.vrvq_flat_tile_order.restype = C.c_int
    for B, T in [(64, 862), (3, 129), (5, 200), (4, 333), (2, 128), (32, 5168), (7, 128), (9, 255), (6, 256), (1, 500), (256, 5168), (17, 131)]:
        n = (B * T + 127) // 128
        order = [L.vrvq_flat_tile_order(B, T, q) for q in range(n)]
        assert sorted(order) == list(range(n)), (B, T)
        assert order[:B - 1] == [(j * T) // 128 for j in range(1, B)], (B, T)
    assert L.vrvq_flat_tile_order(4, 127, 0) == -1 and L.vrvq_flat_tile_order(4, 200, 7) == -1 and L.vrvq_flat_tile_order(4, 200, -1) == -1
