"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the compact code / mask wire format.

Codes: the reference's DACFile stores `codes.numpy().astype(np.uint16)` (models/dac_base.py:34) and reads them back with
`.astype(int)` (:52).  Mask: generate_mask_hard (models/utils.py:55-61) yields a prefix of ones per frame, so a per-frame count
carries it; stages past the count are not payload and are written as 0.  Pinned against the live reference's DACFile in
tests/test_oracle_vs_reference.py (build container only); the product path never imports this module.
"""
import numpy as np


def pack_codes(codes: np.ndarray, mask=None):
    """codes [B,Nq,T] int64, mask [B,Nq,T] float32 0/1 prefix mask or None -> (codes_u16 [B,Nq,T], counts [B,T] uint8 | None)."""
    codes = np.asarray(codes)
    if codes.size and (codes.min() < 0 or codes.max() > 65535):
        raise IndexError("codes outside [0, 65535]")
    u16 = codes.astype(np.uint16)  # models/dac_base.py:34
    if mask is None:
        return u16, None
    mask = np.asarray(mask, dtype=np.float32)
    counts = mask.sum(axis=1)
    k = np.arange(codes.shape[1], dtype=np.float32)[None, :, None]
    if not np.array_equal(mask, (k < counts[:, None, :]).astype(np.float32)):
        raise ValueError("mask is not a 0/1 prefix mask")
    u16 = np.where(k < counts[:, None, :], u16, np.uint16(0)).astype(np.uint16)
    return u16, counts.astype(np.uint8)


def unpack_codes(codes_u16: np.ndarray, counts=None):
    """-> (codes int64 [B,Nq,T], mask float32 [B,Nq,T] | None)."""
    codes = np.asarray(codes_u16).astype(np.int64)  # models/dac_base.py:52 (.astype(int))
    if counts is None:
        return codes, None
    counts = np.asarray(counts)
    if counts.size and counts.max() > codes.shape[1]:
        raise ValueError("count exceeds the number of codebooks")
    k = np.arange(codes.shape[1])[None, :, None]
    return codes, (k < counts[:, None, :]).astype(np.float32)


def payload_bits(counts: np.ndarray, bits_per_codebook) -> int:
    """Bits the kept codes occupy: sum over frames of sum_{k < count} bits[k] (the numerator of cal_bpf_from_mask, utils.py:64-73)."""
    cum = np.concatenate([[0], np.cumsum(np.asarray(bits_per_codebook, dtype=np.int64))])
    return int(cum[np.asarray(counts, dtype=np.int64)].sum())
