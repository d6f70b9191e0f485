"""Compact code / mask wire format on the GPU (csrc/wire.cu through vrvq_pack_codes_u16 / vrvq_unpack_codes_u16): bit-exact
against the numpy oracle, and the encode -> pack -> file -> unpack -> from_codes round trip the format exists for."""
import numpy as np
import pytest
import torch

from oracle import wire as ow
from tests import helpers as H
from tests.golden import gen_inputs as gi
from tests.test_wire_cpu import random_case

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("B,nq,T", [(1, 8, 87), (16, 8, 862), (3, 28, 431), (2, 1, 1), (5, 9, 33), (0, 8, 5), (2, 8, 0), (1, 255, 7)])
def test_pack_unpack_match_the_oracle(B, nq, T):
    from vrvq_b200 import wire

    codes, mask, counts = random_case(11 * B + nq + T, B, nq, T)
    c, m = torch.from_numpy(codes).cuda(), torch.from_numpy(mask).cuda()
    u16, cnt = wire.pack_codes(c, m)
    o_u16, o_cnt = ow.pack_codes(codes, mask)
    assert u16.dtype == torch.uint16 and cnt.dtype == torch.uint8
    assert np.array_equal(npy(u16), o_u16) and np.array_equal(npy(cnt), o_cnt)
    c2, m2 = wire.unpack_codes(u16, cnt)
    o_c2, o_m2 = ow.unpack_codes(o_u16, o_cnt)
    assert c2.dtype == torch.int64 and np.array_equal(npy(c2), o_c2) and np.array_equal(npy(m2), o_m2)
    assert np.array_equal(npy(m2), mask)
    # constant bitrate: no mask, no counts -- the array DACFile.save stores (models/dac_base.py:34)
    u16_all, none = wire.pack_codes(c)
    assert none is None and np.array_equal(npy(u16_all), codes.astype(np.uint16))
    c3, m3 = wire.unpack_codes(u16_all)
    assert m3 is None and np.array_equal(npy(c3), codes)
    if B * T:
        assert wire.payload_bits(cnt, [10] * nq) == ow.payload_bits(counts, [10] * nq) == 10 * int(counts.sum())


def test_views_and_errors():
    from vrvq_b200 import wire

    codes, mask, _ = random_case(3, 4, 8, 200)
    c, m = torch.from_numpy(codes).cuda(), torch.from_numpy(mask).cuda()
    # frame-range and stage-range views (unit stride along T, arbitrary batch / stage strides)
    sl = (slice(1, 4), slice(0, 5), slice(7, 190))
    u16, cnt = wire.pack_codes(c[sl], m[sl])
    o_u16, o_cnt = ow.pack_codes(codes[sl], mask[sl])
    assert np.array_equal(npy(u16), o_u16) and np.array_equal(npy(cnt), o_cnt)
    # a transposed (non unit stride) tensor is made contiguous by the wrapper
    ct = c.permute(0, 2, 1).contiguous().permute(0, 2, 1)
    assert np.array_equal(npy(wire.pack_codes(ct)[0]), codes.astype(np.uint16))
    bad = c.clone()
    bad[2, 3, 100] = 65536
    with pytest.raises(IndexError):
        wire.pack_codes(bad)
    bad[2, 3, 100] = -5
    with pytest.raises(IndexError):
        wire.pack_codes(bad, m)
    holes = m.clone()
    holes[1, 0, 3], holes[1, 1, 3] = 0.0, 1.0
    with pytest.raises(ValueError):
        wire.pack_codes(c, holes)
    with pytest.raises(ValueError):
        wire.pack_codes(c, m * 0.5)
    with pytest.raises(ValueError):
        wire.unpack_codes(wire.pack_codes(c)[0], torch.full((4, 200), 9, dtype=torch.uint8, device="cuda"))


@pytest.mark.parametrize("Nq,B,T,level", [(8, 3, 300, 0.5), (8, 2, 87, 0.25), (12, 2, 131, 1.0)])
def test_encode_pack_file_unpack_decode_round_trip(Nq, B, T, level, tmp_path):
    """What the format is for: the VBR encode's codes + mask survive pack -> .dac file -> load -> unpack, and decoding the
    unpacked codes under the unpacked mask (vrvq_from_codes_f32) reproduces the encoder's z_q; bits on the wire = bpf * frames."""
    import vrvq_b200
    from vrvq_b200 import ops, wire

    sd = gi.torch_state_dict(gi.make_state_dict(900 + Nq, Nq, 1024))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    g = torch.Generator().manual_seed(901 + T)
    z = torch.randn(B, 1024, T, generator=g).cuda()
    imp = torch.rand(B, 1, T, generator=g).cuda()
    out = ops.rvq_encode(pw, z, None, imp, level)
    u16, cnt = wire.pack_codes(out.codes, out.mask)
    assert np.array_equal(npy(cnt), npy(out.mask).sum(1).astype(np.uint8))
    bpf = vrvq_b200.cal_bpf_from_mask(out.mask, [10] * Nq)
    assert wire.payload_bits(cnt, [10] * Nq) == round(bpf * B * T)
    path = wire.DACFile(codes=u16, counts=cnt, chunk_length=T, original_length=512 * T, input_db=torch.tensor([-16.0]), channels=1,
                        sample_rate=44100, padding=True, dac_version="1.0.0").save(tmp_path / "vbr")
    f = wire.DACFile.load(path)
    codes2, mask2 = wire.unpack_codes(f.codes.to(torch.uint16).cuda(), f.counts.cuda())
    assert torch.equal(mask2, out.mask)
    keep = out.mask.bool()
    assert torch.equal(codes2[keep], out.codes[keep]) and not codes2[~keep].any()
    z_q, _, _ = ops.from_codes(pw, codes2, mask2, want_z_p=False)
    H.assert_close_frames(npy(z_q), npy(out.z_q), rtol=1e-5, what="decode(unpack(pack(encode))) vs the encoder's z_q")


def test_streaming_error_flag_defers_the_sync():
    """ADVICE / VERDICT r1: the eager calls read a flag back per call (a sync on a streaming path); with `error_flag=` the
    calls only OR their error bits into a caller-owned device word that is polled once."""
    from vrvq_b200 import ops, wire

    sd = gi.torch_state_dict(gi.make_state_dict(3, 8, 256))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    codes = torch.randint(0, 1024, (2, 8, 50), generator=torch.Generator().manual_seed(1)).cuda()
    mask = vrvq_mask(codes.shape)
    flag = wire.new_error_flag("cuda")
    u16, counts = wire.pack_codes(codes, mask, error_flag=flag)
    c2, m2 = wire.unpack_codes(u16, counts, error_flag=flag)
    zq, _, _ = ops.from_codes(pw, c2, mask=m2, want_z_p=False, error_flag=flag)
    wire.raise_on_flag(flag)  # nothing set
    eager = ops.from_codes(pw, c2, mask=m2, want_z_p=False)[0]
    assert torch.equal(zq, eager)
    bad = codes.clone()
    bad[1, 0, 7] = 5000  # (stage 0 is always kept) fits uint16, outside the codebook: pack accepts it, from_codes flags it
    u16b, _ = wire.pack_codes(bad, mask, error_flag=flag)
    cb, mb = wire.unpack_codes(u16b, counts, error_flag=flag)
    ops.from_codes(pw, cb, mask=None, want_z_p=False, error_flag=flag)
    with pytest.raises(IndexError):
        wire.raise_on_flag(flag)
    flag.zero_()
    notprefix = mask.clone()
    notprefix[0, 2, 0], notprefix[0, 3, 0] = 0.0, 1.0
    wire.pack_codes(codes, notprefix, error_flag=flag)
    with pytest.raises(ValueError):
        wire.raise_on_flag(flag)


def vrvq_mask(shape):
    B, nq, T = shape
    keep = torch.randint(1, nq + 1, (B, 1, T), generator=torch.Generator().manual_seed(2))
    return (torch.arange(nq).view(1, nq, 1) < keep).float().cuda()
