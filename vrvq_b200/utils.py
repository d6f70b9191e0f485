"""Mirrors of the reference mask utilities (models/utils.py:45-73) on the CUDA kernels of libvrvq.so."""
import torch

from . import ops


def generate_mask_hard(x: torch.Tensor, nq: int) -> torch.Tensor:
    """mask[b,k,t] = 1.0 if x[b,0,t] - k >= 0 else 0.0   (models/utils.py:55-61).  x: [B,1,T] CUDA tensor."""
    return ops.generate_mask_hard(x, nq)


def generate_mask_ste(x: torch.Tensor, nq: int, alpha=1) -> torch.Tensor:
    """Forward value of models/utils.py:45-53.  The smooth surrogate only shapes the gradient; its forward value
    equals the hard mask exactly, and this package is inference-only."""
    return ops.generate_mask_hard(x, nq)


def cal_bpf_from_mask(mask: torch.Tensor, bits_per_codebook) -> float:
    """sum(mask * bits) / (B * T) as a python float (models/utils.py:64-73).

    The per-codebook sums are accumulated in binary64 on the device, so unlike the reference's fp32
    torch.sum the result stays exact above 2^24 counted bits."""
    B, nq, T = mask.shape
    if len(bits_per_codebook) != nq:
        raise RuntimeError(f"bits_per_codebook has {len(bits_per_codebook)} entries, mask has {nq} codebooks")
    sums = ops.mask_sums(mask)
    bits = torch.tensor(list(bits_per_codebook), dtype=torch.float64, device=sums.device)
    return float((sums * bits).sum().item() / (B * T))  # .item(): the same device sync the reference has (utils.py:73)


def bpf_from_kept(kept: torch.Tensor, bits_per_codebook, n_frames: int) -> float:
    """Bits per frame from the exact per-stage kept-frame counts the fused encode returns (`kept_frames`)."""
    bits = torch.tensor(list(bits_per_codebook), dtype=torch.float64, device=kept.device)
    return float((kept.to(torch.float64) * bits).sum().item() / n_frames)
