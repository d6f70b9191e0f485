"""Compact code / mask wire format on the GPU, and the reference's `.dac` container around it.

`pack_codes` / `unpack_codes` run the streaming kernels of csrc/wire.cu through the C ABI (`vrvq_pack_codes_u16`,
`vrvq_unpack_codes_u16`): codes as uint16 exactly as the reference's `DACFile.save` stores them (models/dac_base.py:34) and the
hard importance mask (a prefix of ones per frame, models/utils.py:55-61) as one uint8 count per frame -- 2*Nq + 1 bytes per frame
on the wire instead of the 12*Nq of int64 codes + float mask.  `DACFile` mirrors models/dac_base.py:18-58 (same fields, same
`np.save` artifact layout, so files written by either side load on the other) with one optional extra entry, `counts`, which is
what makes the variable bitrate real: the reference has no VBR container (its compress/decompress raise NotImplementedError).
"""
import ctypes as C
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import VrvqError, check, current_stream_ptr

SUPPORTED_VERSIONS = ["1.0.0"]  # models/dac_base.py:15


def new_error_flag(device) -> torch.Tensor:
    """A device-side error word for the streaming form of pack_codes / unpack_codes / ops.from_codes: pass it as
    `error_flag=` and the calls OR their error bits into it WITHOUT synchronising; poll it with `raise_on_flag`
    whenever a sync is convenient (e.g. once per file)."""
    return torch.zeros((1,), dtype=torch.int32, device=device)


def raise_on_flag(flag: torch.Tensor):
    """Read the error word back (one sync) and raise what the eager calls would have raised."""
    f = int(flag.item())
    if f & 1:
        raise IndexError("codes contain a value outside the valid range")
    if f & 2:
        raise ValueError("mask is not a 0/1 prefix mask (generate_mask_hard output) or a frame count exceeds the number of codebooks")


def pack_codes(codes: torch.Tensor, mask: Optional[torch.Tensor] = None, error_flag: Optional[torch.Tensor] = None):
    """codes [B,Nq,T] int64 CUDA (+ mask [B,Nq,T] float32 0/1 prefix mask) -> (codes_u16 [B,Nq,T] uint16, counts [B,T] uint8 | None).

    Stages past a frame's count are not payload and come out as 0.  Raises IndexError for codes outside [0, 65535] and
    ValueError for a mask that is not a prefix of ones (one device read-back, like from_codes) -- unless the caller passes
    `error_flag` (see new_error_flag): then nothing synchronises and the caller polls the flag."""
    if not codes.is_cuda or codes.dtype != torch.int64 or codes.dim() != 3:
        raise VrvqError("codes must be a CUDA int64 tensor [B, Nq, T] (no CPU fallback)")
    B, nq, T = codes.shape
    if nq > 255:
        raise VrvqError("at most 255 codebooks fit the uint8 count")
    if codes.stride(-1) != 1 and T > 1:
        codes = codes.contiguous()
    dev = codes.device
    out = torch.empty((B, nq, T), dtype=torch.uint16, device=dev)
    counts = None
    m_ptr, m_sb, m_sq = None, 0, 0
    if mask is not None:
        if not mask.is_cuda or mask.dtype != torch.float32 or tuple(mask.shape) != (B, nq, T):
            raise VrvqError("mask must be a CUDA float32 tensor [B, Nq, T]")
        if mask.stride(-1) != 1 and T > 1:
            mask = mask.contiguous()
        counts = torch.empty((B, T), dtype=torch.uint8, device=dev)
        m_ptr, m_sb, m_sq = mask.data_ptr(), mask.stride(0), mask.stride(1)
    if out.numel() == 0:
        return out, counts
    flag = error_flag if error_flag is not None else torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().vrvq_pack_codes_u16(codes.data_ptr(), codes.stride(0), codes.stride(1), m_ptr, m_sb, m_sq, B, T, nq, out.data_ptr(),
                                             counts.data_ptr() if counts is not None else None, flag.data_ptr(), current_stream_ptr(dev)),
              "vrvq_pack_codes_u16")
    _lib.count_launch()
    if error_flag is not None:
        return out, counts
    f = int(flag.item())
    if f & 1:
        raise IndexError("codes contain a value outside [0, 65535]")
    if f & 2:
        raise ValueError("mask is not a 0/1 prefix mask (generate_mask_hard output)")
    return out, counts


def unpack_codes(codes_u16: torch.Tensor, counts: Optional[torch.Tensor] = None, error_flag: Optional[torch.Tensor] = None):
    """(codes_u16 [B,Nq,T] uint16, counts [B,T] uint8 | None) CUDA -> (codes int64 [B,Nq,T], mask float32 [B,Nq,T] | None).
    `error_flag`: as in pack_codes (no synchronisation; the caller polls)."""
    if not codes_u16.is_cuda or codes_u16.dtype != torch.uint16 or codes_u16.dim() != 3:
        raise VrvqError("codes_u16 must be a CUDA uint16 tensor [B, Nq, T] (no CPU fallback)")
    codes_u16 = codes_u16.contiguous()
    B, nq, T = codes_u16.shape
    if nq > 255:
        raise VrvqError("at most 255 codebooks fit the uint8 count")
    dev = codes_u16.device
    codes = torch.empty((B, nq, T), dtype=torch.int64, device=dev)
    mask = None
    if counts is not None:
        if not counts.is_cuda or counts.dtype != torch.uint8 or tuple(counts.shape) != (B, T):
            raise VrvqError("counts must be a CUDA uint8 tensor [B, T]")
        counts = counts.contiguous()
        mask = torch.empty((B, nq, T), dtype=torch.float32, device=dev)
    if codes.numel() == 0:
        return codes, mask
    flag = error_flag if error_flag is not None else torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().vrvq_unpack_codes_u16(codes_u16.data_ptr(), counts.data_ptr() if counts is not None else None, B, T, nq,
                                               codes.data_ptr(), codes.stride(0), codes.stride(1),
                                               mask.data_ptr() if mask is not None else None, mask.stride(0) if mask is not None else 0,
                                               mask.stride(1) if mask is not None else 0, flag.data_ptr(), current_stream_ptr(dev)),
              "vrvq_unpack_codes_u16")
    _lib.count_launch()
    if error_flag is None and int(flag.item()) & 2:
        raise ValueError("a frame count exceeds the number of codebooks")
    return codes, mask


def payload_bits(counts: torch.Tensor, bits_per_codebook) -> int:
    """Bits the kept codes occupy (the numerator of cal_bpf_from_mask, models/utils.py:64-73), exact in int64."""
    cum = torch.zeros(len(bits_per_codebook) + 1, dtype=torch.int64, device=counts.device)
    cum[1:] = torch.cumsum(torch.tensor(list(bits_per_codebook), dtype=torch.int64, device=counts.device), 0)
    return int(cum[counts.to(torch.int64)].sum().item())


@dataclass
class DACFile:
    """models/dac_base.py:18-58 with the optional per-frame `counts` of the variable-bitrate mask."""

    codes: torch.Tensor

    # Metadata
    chunk_length: int
    original_length: int
    input_db: float
    channels: int
    sample_rate: int
    padding: bool
    dac_version: str
    counts: Optional[torch.Tensor] = None

    def save(self, path):
        codes = self.codes.detach().cpu()
        codes = codes.numpy() if codes.dtype == torch.uint16 else codes.numpy().astype(np.uint16)
        input_db = self.input_db.detach().cpu().numpy() if isinstance(self.input_db, torch.Tensor) else np.asarray(self.input_db)
        artifacts = {
            "codes": codes,
            "metadata": {
                "input_db": input_db.astype(np.float32),
                "original_length": self.original_length,
                "sample_rate": self.sample_rate,
                "chunk_length": self.chunk_length,
                "channels": self.channels,
                "padding": self.padding,
                "dac_version": SUPPORTED_VERSIONS[-1],
            },
        }
        if self.counts is not None:
            artifacts["counts"] = self.counts.detach().cpu().numpy().astype(np.uint8)
        path = Path(path).with_suffix(".dac")
        with open(path, "wb") as f:
            np.save(f, artifacts)
        return path

    @classmethod
    def load(cls, path):
        artifacts = np.load(path, allow_pickle=True)[()]
        codes = torch.from_numpy(artifacts["codes"].astype(int))
        if artifacts["metadata"].get("dac_version", None) not in SUPPORTED_VERSIONS:
            raise RuntimeError(f"Given file {path} can't be loaded with this version of descript-audio-codec.")
        counts = artifacts.get("counts", None)
        return cls(codes=codes, counts=None if counts is None else torch.from_numpy(counts), **artifacts["metadata"])
