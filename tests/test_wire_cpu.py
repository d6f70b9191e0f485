"""Compact code / mask wire format: the numpy oracle, the `.dac` container mirror and the C-ABI
argument checks -- everything that runs without a GPU."""
import numpy as np
import pytest
import torch

import vrvq_b200
from oracle import wire as ow
from vrvq_b200 import _lib, wire


def random_case(seed, B, nq, T):
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 1024, size=(B, nq, T), dtype=np.int64)
    counts = rng.integers(1, nq + 1, size=(B, T))  # mask_imp[:,0,:] is always 1
    mask = (np.arange(nq)[None, :, None] < counts[:, None, :]).astype(np.float32)
    return codes, mask, counts


@pytest.mark.parametrize("B,nq,T", [(1, 8, 87), (3, 28, 41), (2, 1, 1), (0, 8, 5), (2, 8, 0)])
def test_oracle_round_trip(B, nq, T):
    codes, mask, counts = random_case(B * 100 + nq + T, B, nq, T)
    u16, cnt = ow.pack_codes(codes, mask)
    assert u16.dtype == np.uint16 and cnt.dtype == np.uint8 and cnt.shape == (B, T)
    assert np.array_equal(cnt, counts)
    kept = mask.astype(bool)
    assert np.array_equal(u16[kept], codes[kept].astype(np.uint16)) and not u16[~kept].any()
    c2, m2 = ow.unpack_codes(u16, cnt)
    assert c2.dtype == np.int64 and np.array_equal(m2, mask) and np.array_equal(c2[kept], codes[kept])
    # without a mask the packed codes are exactly what the reference's DACFile.save stores (dac_base.py:34)
    u16_all, none = ow.pack_codes(codes)
    assert none is None and np.array_equal(u16_all, codes.astype(np.uint16))
    assert np.array_equal(ow.unpack_codes(u16_all)[0], codes)
    bits = [10] * nq
    assert ow.payload_bits(cnt, bits) == int((mask * np.asarray(bits, dtype=np.float64)[None, :, None]).sum())


def test_oracle_rejects_what_the_format_cannot_carry():
    codes, mask, _ = random_case(5, 2, 8, 16)
    bad = codes.copy()
    bad[0, 0, 0] = 65536
    with pytest.raises(IndexError):
        ow.pack_codes(bad)
    bad[0, 0, 0] = -1
    with pytest.raises(IndexError):
        ow.pack_codes(bad)
    holes = mask.copy()
    holes[1, 0, 3], holes[1, 1, 3] = 0.0, 1.0
    with pytest.raises(ValueError):
        ow.pack_codes(codes, holes)
    with pytest.raises(ValueError):
        ow.pack_codes(codes, mask * 0.5)
    with pytest.raises(ValueError):
        ow.unpack_codes(codes.astype(np.uint16), np.full((2, 16), 9, dtype=np.uint8))


def test_dac_file_round_trip_and_reference_layout(tmp_path):
    codes, mask, counts = random_case(7, 2, 8, 50)
    u16, cnt = ow.pack_codes(codes, mask)
    meta = dict(chunk_length=50, original_length=25600, input_db=torch.tensor([-16.0, -20.5]), channels=1, sample_rate=44100, padding=True,
                dac_version="1.0.0")
    f = wire.DACFile(codes=torch.from_numpy(u16), counts=torch.from_numpy(cnt), **meta)
    path = f.save(tmp_path / "clip")
    assert path.suffix == ".dac"
    g = vrvq_b200.DACFile.load(path)
    assert g.codes.dtype == torch.int64 and np.array_equal(g.codes.numpy(), u16.astype(np.int64))
    assert g.counts.dtype == torch.uint8 and np.array_equal(g.counts.numpy(), cnt)
    assert (g.chunk_length, g.original_length, g.channels, g.sample_rate, g.padding, g.dac_version) == (50, 25600, 1, 44100, True, "1.0.0")
    assert np.array_equal(g.input_db, np.asarray([-16.0, -20.5], dtype=np.float32))
    # int64 codes are narrowed on save exactly like the reference does; without counts the artifact has the reference's two keys
    path2 = wire.DACFile(codes=torch.from_numpy(codes), **meta).save(tmp_path / "cbr.dac")
    raw = np.load(path2, allow_pickle=True)[()]
    assert set(raw) == {"codes", "metadata"} and raw["codes"].dtype == np.uint16 and np.array_equal(raw["codes"], codes.astype(np.uint16))
    assert set(raw["metadata"]) == {"input_db", "original_length", "sample_rate", "chunk_length", "channels", "padding", "dac_version"}
    # a file laid out the way models/dac_base.py:32-48 writes it loads here
    ref_style = {"codes": codes.astype(np.uint16), "metadata": {"input_db": np.float32([-3.0]), "original_length": 1, "sample_rate": 16000,
                                                                 "chunk_length": 2, "channels": 1, "padding": False, "dac_version": "1.0.0"}}
    with open(tmp_path / "ref.dac", "wb") as fh:
        np.save(fh, ref_style)
    h = wire.DACFile.load(tmp_path / "ref.dac")
    assert h.counts is None and np.array_equal(h.codes.numpy(), codes) and h.sample_rate == 16000
    ref_style["metadata"]["dac_version"] = "0.9"
    with open(tmp_path / "old.dac", "wb") as fh:
        np.save(fh, ref_style)
    with pytest.raises(RuntimeError):
        wire.DACFile.load(tmp_path / "old.dac")


def test_abi_argument_checks_without_a_gpu():
    L = _lib.lib()
    assert L.vrvq_pack_codes_u16(None, 0, 0, None, 0, 0, 1, 4, 8, None, None, None, None) == -1
    assert L.vrvq_unpack_codes_u16(None, None, 1, 4, 8, None, 0, 0, None, 0, 0, None, None) == -1
    assert L.vrvq_pack_codes_u16(None, 0, 0, None, 0, 0, -1, 4, 8, None, None, None, None) == -1
    c = np.zeros((1, 8, 4), dtype=np.int64)
    o = np.zeros((1, 8, 4), dtype=np.uint16)
    m = np.ones((1, 8, 4), dtype=np.float32)
    # a mask without a counts buffer, and more codebooks than a uint8 count can carry
    assert L.vrvq_pack_codes_u16(c.ctypes.data, 32, 4, m.ctypes.data, 32, 4, 1, 4, 8, o.ctypes.data, None, None, None) == -1
    assert L.vrvq_pack_codes_u16(c.ctypes.data, 32, 4, None, 0, 0, 1, 4, 256, o.ctypes.data, None, None, None) == -1
    assert b"vrvq_pack_codes_u16" in L.vrvq_last_error()
    if not torch.cuda.is_available():
        assert L.vrvq_pack_codes_u16(c.ctypes.data, 32, 4, None, 0, 0, 1, 4, 8, o.ctypes.data, None, None, None) == -4
        assert L.vrvq_unpack_codes_u16(o.ctypes.data, None, 1, 4, 8, c.ctypes.data, 32, 4, None, 0, 0, None, None) == -4
    # nothing to do: ok without touching a device
    assert L.vrvq_pack_codes_u16(None, 0, 0, None, 0, 0, 0, 4, 8, None, None, None, None) in (0, -4)


def test_no_cpu_fallback():
    with pytest.raises(vrvq_b200.VrvqError):
        wire.pack_codes(torch.zeros(1, 8, 4, dtype=torch.int64))
    with pytest.raises(vrvq_b200.VrvqError):
        wire.unpack_codes(torch.zeros(1, 8, 4, dtype=torch.uint16))
