"""Functional front-end of the C ABI: packed weights and the fused launches, on torch CUDA tensors.

These are the calls the nn.Module mirrors in quantize.py / utils.py are built on; bench.py times
`rvq_encode_into` directly (device-resident inputs, pre-allocated outputs, one kernel per call).
"""
import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import CD, EncodeArgs, FromCodesArgs, VrvqError, check, current_stream_ptr, ptr, require_cuda_f32

TILE_FRAMES = 32  # shard/padding granularity in frames (the CUDA-core kernel's tile; the tensor-core kernel picks multiples of 8)


def fold_weight_norm(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """Effective weight of a weight-normed conv: torch._weight_norm(v, g, 0) evaluated on the CPU in fp32,
    which is bit-identical to what the reference's weight_norm hook computes on its CPU path
    (models/layers.py:17-18)."""
    return torch._weight_norm(v.detach().to("cpu", torch.float32), g.detach().to("cpu", torch.float32), 0)


class PackedWeights:
    """Device-resident weight blob of an Nq-stage RVQ (layout: csrc/common.cuh)."""

    def __init__(self, w_in, b_in, w_out, b_out, codebook, device):
        # all CPU float32: w_in [Nq,8,D], b_in [Nq,8], w_out [Nq,D,8], b_out [Nq,D], codebook [Nq,K,8]
        L = _lib.lib()
        arrs = [np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy()) for t in (w_in, b_in, w_out, b_out, codebook)]
        self.n_codebooks, cd, self.input_dim = arrs[0].shape
        self.codebook_size = arrs[4].shape[1]
        if cd != CD or arrs[4].shape[2] != CD:
            raise VrvqError(f"codebook_dim must be {CD} (conf/base.yml:11); got {cd}")
        if arrs[2].shape != (self.n_codebooks, self.input_dim, CD) or arrs[1].shape != (self.n_codebooks, CD) \
                or arrs[3].shape != (self.n_codebooks, self.input_dim) or arrs[4].shape[0] != self.n_codebooks:
            raise VrvqError("inconsistent weight shapes")
        nbytes = L.vrvq_blob_bytes(self.n_codebooks, self.input_dim, self.codebook_size, CD)
        if nbytes == 0:
            raise VrvqError("vrvq_blob_bytes rejected the shape")
        host = np.empty(nbytes // 4, np.float32)
        check(L.vrvq_pack_weights(self.n_codebooks, self.input_dim, self.codebook_size, CD, *[a.ctypes.data for a in arrs],
                                  host.ctypes.data, nbytes), "vrvq_pack_weights")
        self.host_blob = host
        self.device = torch.device(device)
        self.blob = torch.from_numpy(host).to(self.device) if self.device.type == "cuda" else None

    def supported(self) -> bool:
        return bool(_lib.lib().vrvq_supported(self.input_dim, self.codebook_size, CD))

    def normalized_codebook(self, stage: int):
        """(F.normalize(codebook), sum of squares) as stored in the blob -- host read-back for tests."""
        cb = np.empty((self.codebook_size, CD), np.float32)
        c2 = np.empty((self.codebook_size,), np.float32)
        check(_lib.lib().vrvq_blob_codebook(self.host_blob.ctypes.data, self.host_blob.nbytes, stage, cb.ctypes.data, c2.ctypes.data),
              "vrvq_blob_codebook")
        return cb, c2

    @classmethod
    def from_state_dict(cls, sd, device, prefix=""):
        """Reference layout: {prefix}quantizers.{i}.in_proj.weight_g/_v/bias, .out_proj.*, .codebook.weight."""
        n = 0
        while f"{prefix}quantizers.{n}.codebook.weight" in sd:
            n += 1
        if n == 0:
            raise KeyError(f"no '{prefix}quantizers.0.codebook.weight' in state dict")
        w_in, b_in, w_out, b_out, cb = [], [], [], [], []
        for i in range(n):
            q = f"{prefix}quantizers.{i}."
            w_in.append(fold_weight_norm(sd[q + "in_proj.weight_v"], sd[q + "in_proj.weight_g"])[:, :, 0])
            b_in.append(sd[q + "in_proj.bias"].detach().to("cpu", torch.float32))
            w_out.append(fold_weight_norm(sd[q + "out_proj.weight_v"], sd[q + "out_proj.weight_g"])[:, :, 0])
            b_out.append(sd[q + "out_proj.bias"].detach().to("cpu", torch.float32))
            cb.append(sd[q + "codebook.weight"].detach().to("cpu", torch.float32))
        return cls(torch.stack(w_in), torch.stack(b_in), torch.stack(w_out), torch.stack(b_out), torch.stack(cb), device)


def _check_view(t: torch.Tensor, name: str):
    if t.stride(-1) != 1 and t.shape[-1] > 1:
        raise VrvqError(f"{name} must have unit stride along T (got strides {t.stride()})")


class EncodeOutputs:
    """Pre-allocated outputs + accumulators of one encode call (re-usable across calls of equal shape)."""

    def __init__(self, B, D, T, n_run, device, z_q=True, z_q_is=False, latents=True, mask=True, loss_pf=False, pad_z_q_is_rows=False):
        f32 = dict(dtype=torch.float32, device=device)
        self.codes = torch.empty((B, n_run, T), dtype=torch.int64, device=device)
        self.z_q = torch.empty((B, D, T), **f32) if z_q else None
        if z_q_is and pad_z_q_is_rows and T % TILE_FRAMES:
            # opt-in: row pitch rounded up to a tile (128 bytes), returned as a [..., :T] view.  Every row then starts on a
            # cache line, which removes the partial-sector stores of a T like 862 ("Stores").
            tp = (T + TILE_FRAMES - 1) // TILE_FRAMES * TILE_FRAMES
            self.z_q_is = torch.empty((B, n_run, D, tp), **f32)[..., :T]
        else:
            self.z_q_is = torch.empty((B, n_run, D, T), **f32) if z_q_is else None
        self.latents = torch.empty((B, CD * n_run, T), **f32) if latents else None
        self.mask = torch.empty((B, n_run, T), **f32) if mask else None
        self.loss_pf = torch.empty((B, n_run, T), **f32) if loss_pf else None
        # one int64 buffer: [0] = masked loss sum (binary64 bits), [8:8+n_run] = kept-frame counts
        self.accum = torch.zeros((8 + 32,), dtype=torch.int64, device=device)
        self.n_run = n_run
        self.frames = B * T

    @property
    def loss_sum(self):
        return self.accum[:1].view(torch.float64)

    @property
    def kept(self):
        return self.accum[8:8 + self.n_run]


def rvq_encode_into(w: PackedWeights, z: torch.Tensor, out: EncodeOutputs, n_run: int, imp_map: Optional[torch.Tensor] = None,
                    level=None, zero_accum: bool = True, stream=None):
    """One fused launch (vrvq_rvq_encode_f32).  z [B,D,T] fp32 CUDA (any batch/row stride, unit stride along T);
    imp_map [B,1,T] or [B,T] or None (CBR masking); level: python number or CUDA tensor with 1 or B elements."""
    require_cuda_f32(z, "z")
    if z.dim() != 3 or z.shape[1] != w.input_dim:
        raise VrvqError(f"z must be [B, {w.input_dim}, T], got {tuple(z.shape)}")
    if w.blob is None or w.blob.device != z.device:
        raise VrvqError(f"weights are on {w.device}, z on {z.device}")
    _check_view(z, "z")
    B, D, T = z.shape
    a = EncodeArgs()
    a.struct_size = C.sizeof(EncodeArgs)
    a.B, a.T, a.input_dim, a.n_codebooks, a.codebook_size, a.n_run = B, T, D, w.n_codebooks, w.codebook_size, int(n_run)
    a.blob = w.blob.data_ptr()
    a.z, a.z_stride_b, a.z_stride_d = z.data_ptr(), z.stride(0), z.stride(1)
    keep = [z]
    if imp_map is not None:
        require_cuda_f32(imp_map, "imp_map")
        imp2 = imp_map.reshape(B, T) if imp_map.dim() == 3 else imp_map
        if imp2.shape != (B, T):
            raise VrvqError(f"imp_map must be [B,1,T]/[B,T], got {tuple(imp_map.shape)}")
        _check_view(imp2, "imp_map")
        a.imp_map, a.imp_stride_b = imp2.data_ptr(), imp2.stride(0)
        keep.append(imp2)
        if level is None:
            raise VrvqError("level must be given with imp_map (quantize.py:348)")
        if isinstance(level, torch.Tensor):
            require_cuda_f32(level, "level")
            lv = level.reshape(-1)
            if lv.numel() not in (1, B):
                raise VrvqError("level tensor must have 1 or B elements")
            a.level_dev, a.level_stride = lv.data_ptr(), (0 if lv.numel() == 1 else lv.stride(0))
            keep.append(lv)
        else:
            a.level_dev, a.level_host = None, float(level)

    def bind(t, names, what):
        if t is None:
            return
        _check_view(t, what)
        setattr(a, names[0], t.data_ptr())
        for n, s in zip(names[1:], t.stride()[:-1]):
            setattr(a, n, s)

    bind(out.codes, ("codes", "codes_stride_b", "codes_stride_q"), "codes")
    bind(out.z_q, ("z_q", "z_q_stride_b", "z_q_stride_d"), "z_q")
    bind(out.z_q_is, ("z_q_is", "z_q_is_stride_b", "z_q_is_stride_q", "z_q_is_stride_d"), "z_q_is")
    bind(out.latents, ("latents", "latents_stride_b", "latents_stride_c"), "latents")
    bind(out.mask, ("mask", "mask_stride_b", "mask_stride_q"), "mask")
    bind(out.loss_pf, ("loss_pf", "loss_pf_stride_b", "loss_pf_stride_q"), "loss_pf")
    if zero_accum:
        out.accum.zero_()
    a.loss_masked_sum = out.accum.data_ptr()
    a.kept = out.accum.data_ptr() + 8 * 8
    st = current_stream_ptr(z.device) if stream is None else C.c_void_p(stream)
    with torch.cuda.device(z.device):
        check(_lib.lib().vrvq_rvq_encode_f32(C.byref(a), st), "vrvq_rvq_encode_f32")
    _lib.count_launch()
    return out


def encode_launch_info(w: PackedWeights, B: int, T: int, n_run: int, device, z_q_is: bool = False):
    a = EncodeArgs()
    a.struct_size = C.sizeof(EncodeArgs)
    a.B, a.T, a.input_dim, a.n_codebooks, a.codebook_size, a.n_run = B, T, w.input_dim, w.n_codebooks, w.codebook_size, n_run
    dummy = w.blob.data_ptr()
    a.blob, a.z, a.codes = dummy, dummy, dummy
    a.z_stride_b, a.z_stride_d = w.input_dim * T, T
    if z_q_is:  # the per-stage outputs change the kernel choice and the tiling
        a.z_q_is = dummy
        a.z_q_is_stride_b, a.z_q_is_stride_q, a.z_q_is_stride_d = n_run * w.input_dim * T, w.input_dim * T, T
    g, b, s = C.c_int(), C.c_int(), C.c_int()
    with torch.cuda.device(device):
        check(_lib.lib().vrvq_rvq_encode_launch_info(C.byref(a), C.byref(g), C.byref(b), C.byref(s)), "vrvq_rvq_encode_launch_info")
        name = _lib.lib().vrvq_rvq_encode_kernel_name(C.byref(a))
    return {"grid": g.value, "block": b.value, "smem_bytes": s.value, "kernel": name.decode() if name else None}


def tc_kernel_available(w: PackedWeights, n_run: int, z_q_is: bool = False) -> bool:
    """Whether a tensor-core (tcgen05) instantiation of the fused encode exists for this model / call shape: D in
    {256, 512, 1024}, K = 1024, at most 32 codebooks -- and at most 8 when the per-stage outputs z_q_is are requested
    (csrc/common.cuh: tc_shape_ok, TC_MAX_NQ_ZQIS)."""
    return w.codebook_size == 1024 and w.input_dim in (256, 512, 1024) and 1 <= w.n_codebooks <= (8 if z_q_is else 32)


def rvq_encode(w: PackedWeights, z, n_run=None, imp_map=None, level=None, want_z_q_is=False, want_loss_pf=False):
    """Allocate outputs and run the fused encode.  Returns the EncodeOutputs."""
    n_run = w.n_codebooks if n_run is None else int(n_run)
    B, D, T = z.shape
    out = EncodeOutputs(B, D, T, n_run, z.device, z_q=True, z_q_is=want_z_q_is, latents=True, mask=True, loss_pf=want_loss_pf)
    if B * T == 0:
        return out
    return rvq_encode_into(w, z, out, n_run, imp_map, level, zero_accum=False)


def from_codes(w: PackedWeights, codes: torch.Tensor, mask: Optional[torch.Tensor] = None, want_z_q_is=False, want_z_p=True,
               error_flag: Optional[torch.Tensor] = None):
    """vrvq_from_codes_f32: codes [B,n,T] int64 CUDA -> (z_q [B,D,T], z_p [B,8n,T] | None, z_q_is | None).
    An out-of-range code raises IndexError as F.embedding does (one flag read-back = one sync per call); a streaming caller
    passes `error_flag` (wire.new_error_flag) instead: the kernel sets bit 0 there, nothing synchronises, and the caller polls
    it with wire.raise_on_flag when convenient."""
    if not codes.is_cuda or codes.dtype != torch.int64 or codes.dim() != 3:
        raise VrvqError("codes must be a CUDA int64 tensor [B, n, T] (no CPU fallback)")
    if w.blob is None or w.blob.device != codes.device:
        raise VrvqError(f"weights are on {w.device}, codes on {codes.device}")
    _check_view(codes, "codes")
    B, n, T = codes.shape
    dev = codes.device
    z_q = torch.empty((B, w.input_dim, T), dtype=torch.float32, device=dev)
    z_p = torch.empty((B, CD * n, T), dtype=torch.float32, device=dev) if want_z_p else None
    z_q_is = torch.empty((B, n, w.input_dim, T), dtype=torch.float32, device=dev) if want_z_q_is else None
    if B * T == 0:
        return z_q, z_p, z_q_is
    flag = error_flag if error_flag is not None else torch.zeros((1,), dtype=torch.int32, device=dev)
    a = FromCodesArgs()
    a.struct_size = C.sizeof(FromCodesArgs)
    a.B, a.T, a.input_dim, a.n_codebooks, a.codebook_size, a.n_run = B, T, w.input_dim, w.n_codebooks, w.codebook_size, n
    a.blob = w.blob.data_ptr()
    a.codes, a.codes_stride_b, a.codes_stride_q = codes.data_ptr(), codes.stride(0), codes.stride(1)
    if mask is not None:
        require_cuda_f32(mask, "mask")
        if mask.shape != (B, n, T):
            raise VrvqError("mask must be [B, n, T]")
        _check_view(mask, "mask")
        a.mask, a.mask_stride_b, a.mask_stride_q = mask.data_ptr(), mask.stride(0), mask.stride(1)
    a.z_q, a.z_q_stride_b, a.z_q_stride_d = z_q.data_ptr(), z_q.stride(0), z_q.stride(1)
    if z_p is not None:
        a.z_p, a.z_p_stride_b, a.z_p_stride_c = z_p.data_ptr(), z_p.stride(0), z_p.stride(1)
    if z_q_is is not None:
        a.z_q_is = z_q_is.data_ptr()
        a.z_q_is_stride_b, a.z_q_is_stride_q, a.z_q_is_stride_d = z_q_is.stride(0), z_q_is.stride(1), z_q_is.stride(2)
    a.error_flag = flag.data_ptr()
    with torch.cuda.device(dev):
        check(_lib.lib().vrvq_from_codes_f32(C.byref(a), current_stream_ptr(dev)), "vrvq_from_codes_f32")
    _lib.count_launch()
    if error_flag is None and int(flag.item()) != 0:  # F.embedding raises on out-of-range indices (quantize.py:82)
        raise IndexError("codes contain an index outside [0, codebook_size)")
    return z_q, z_p, z_q_is


def search_latents(w: PackedWeights, latents: torch.Tensor) -> torch.Tensor:
    """vrvq_search_latents_f32: latents [B, 8n, T] fp32 CUDA -> codes [B, n, T] int64."""
    require_cuda_f32(latents, "latents")
    if latents.dim() != 3 or latents.shape[1] % CD != 0:
        raise VrvqError(f"latents must be [B, 8n, T], got {tuple(latents.shape)}")
    if w.blob is None or w.blob.device != latents.device:
        raise VrvqError(f"weights are on {w.device}, latents on {latents.device}")
    _check_view(latents, "latents")
    B, c, T = latents.shape
    n = min(c // CD, w.n_codebooks)  # quantize.py:271-275
    codes = torch.empty((B, n, T), dtype=torch.int64, device=latents.device)
    if codes.numel() == 0:
        return codes
    with torch.cuda.device(latents.device):
        check(_lib.lib().vrvq_search_latents_f32(w.blob.data_ptr(), w.n_codebooks, w.input_dim, w.codebook_size, latents.data_ptr(),
                                                 latents.stride(0), latents.stride(1), B, T, n, codes.data_ptr(), codes.stride(0),
                                                 codes.stride(1), current_stream_ptr(latents.device)), "vrvq_search_latents_f32")
    _lib.count_launch()
    return codes


def generate_mask_hard(x: torch.Tensor, nq: int) -> torch.Tensor:
    """vrvq_generate_mask_hard_f32; x [B,1,T] (float32; integer tensors are converted as the reference's type promotion does)."""
    if not x.is_cuda:
        raise VrvqError("generate_mask_hard: x must be a CUDA tensor (no CPU fallback)")
    if x.dim() != 3 or x.shape[1] != 1:
        raise VrvqError(f"x must be [B,1,T], got {tuple(x.shape)}")
    xf = x.to(torch.float32).contiguous()
    B, _, T = xf.shape
    mask = torch.empty((B, int(nq), T), dtype=torch.float32, device=x.device)
    if mask.numel() == 0:
        return mask
    with torch.cuda.device(x.device):
        check(_lib.lib().vrvq_generate_mask_hard_f32(xf.data_ptr(), xf.stride(0), B, T, int(nq), mask.data_ptr(), mask.stride(0),
                                                     mask.stride(1), current_stream_ptr(x.device)), "vrvq_generate_mask_hard_f32")
    _lib.count_launch()
    return mask


def mask_sums(mask: torch.Tensor) -> torch.Tensor:
    """vrvq_mask_sum_f32: per-codebook sum over (b,t) in binary64, [nq] CUDA tensor (no sync)."""
    require_cuda_f32(mask, "mask")
    if mask.dim() != 3:
        raise VrvqError("mask must be [B, Nq, T]")
    m = mask if mask.stride(-1) == 1 else mask.contiguous()
    B, nq, T = m.shape
    sums = torch.zeros((nq,), dtype=torch.float64, device=m.device)
    if m.numel() == 0:
        return sums
    with torch.cuda.device(m.device):
        check(_lib.lib().vrvq_mask_sum_f32(m.data_ptr(), m.stride(0), m.stride(1), B, T, nq, sums.data_ptr(), current_stream_ptr(m.device)),
              "vrvq_mask_sum_f32")
    _lib.count_launch()
    return sums


def remask(z_q_is: torch.Tensor, imp_map: torch.Tensor, level_scaled: float, want_mask=True):
    """vrvq_remask_f32 -- one level of the sweep in scripts/inference.py:95-100.
    Returns (z_q [B,D,T], mask [B,Nq,T] | None, kept [Nq] int64 CUDA)."""
    require_cuda_f32(z_q_is, "z_q_is")
    require_cuda_f32(imp_map, "imp_map")
    if z_q_is.dim() != 4:
        raise VrvqError("z_q_is must be [B, Nq, D, T]")
    _check_view(z_q_is, "z_q_is")
    B, nq, D, T = z_q_is.shape
    imp2 = imp_map.reshape(B, T)
    _check_view(imp2, "imp_map")
    dev = z_q_is.device
    z_q = torch.empty((B, D, T), dtype=torch.float32, device=dev)
    mask = torch.empty((B, nq, T), dtype=torch.float32, device=dev) if want_mask else None
    kept = torch.zeros((nq,), dtype=torch.int64, device=dev)
    if B * T * D == 0:
        return z_q, mask, kept
    with torch.cuda.device(dev):
        check(_lib.lib().vrvq_remask_f32(z_q_is.data_ptr(), z_q_is.stride(0), z_q_is.stride(1), z_q_is.stride(2), imp2.data_ptr(),
                                         imp2.stride(0), C.c_float(float(level_scaled)), B, D, T, nq, z_q.data_ptr(), z_q.stride(0),
                                         z_q.stride(1), ptr(mask), mask.stride(0) if want_mask else 0, mask.stride(1) if want_mask else 0,
                                         kept.data_ptr(), current_stream_ptr(dev)), "vrvq_remask_f32")
    _lib.count_launch()
    return z_q, mask, kept


class PackedConv3:
    """Device-resident weights of one `Snake1d -> WNConv1d(k=3, padding=1)` block of the importance subnet
    (models/importance_subnet.py:18-34): alpha [Cin], packed conv weight [Cin*3][Cout_padded], bias [Cout]."""

    def __init__(self, alpha, weight, bias, device):
        # CPU float32: alpha [Cin] (Snake1d.alpha flattened), weight [Cout,Cin,3] already weight-norm folded, bias [Cout]
        L = _lib.lib()
        weight = weight.detach().to("cpu", torch.float32).contiguous()
        if weight.dim() != 3 or weight.shape[2] != 3:
            raise VrvqError(f"conv weight must be [Cout, Cin, 3], got {tuple(weight.shape)}")
        self.cout, self.cin = int(weight.shape[0]), int(weight.shape[1])
        if self.cin % 8 != 0:
            raise VrvqError(f"importance-subnet widths must be multiples of 8 (got Cin={self.cin})")
        n = L.vrvq_conv3_packed_floats(self.cout, self.cin)
        packed = torch.empty(n, dtype=torch.float32)
        check(L.vrvq_pack_conv3_weights(self.cout, self.cin, weight.data_ptr(), packed.data_ptr(), n), "vrvq_pack_conv3_weights")
        self.packed = packed.to(torch.device(device))
        self.device = self.packed.device  # resolved ("cuda" -> "cuda:0")
        # the wide blocks (Cin % 64 == 0, Cout % 128 == 0) also get the tensor-core operand tiles (csrc/subnet_tc.cu)
        self.packed_tc = None
        n_tc = L.vrvq_conv3_tc_packed_floats(self.cout, self.cin)
        if n_tc:
            ptc = torch.empty(n_tc, dtype=torch.float32)
            check(L.vrvq_pack_conv3_tc_weights(self.cout, self.cin, weight.data_ptr(), ptc.data_ptr(), n_tc), "vrvq_pack_conv3_tc_weights")
            self.packed_tc = ptc.to(self.device)
        self.alpha = alpha.detach().to("cpu", torch.float32).reshape(-1).contiguous().to(self.device)
        self.bias = bias.detach().to("cpu", torch.float32).reshape(-1).contiguous().to(self.device)
        if self.alpha.numel() != self.cin or self.bias.numel() != self.cout:
            raise VrvqError("alpha must have Cin entries and bias Cout entries")


def _tc_block_ok(w: PackedConv3, x: torch.Tensor) -> bool:
    import os

    return w.packed_tc is not None and os.environ.get("VRVQ_SUBNET_IMPL", "") != "cuda" and (x.stride(0) % 4 == 0 or x.shape[0] == 1)


def _padded_rows(B: int, Cc: int, T: int, device) -> torch.Tensor:
    """[B, C, T] view of a buffer whose rows are padded to a multiple of 4 frames (16 bytes): the tensor-core block stores such an
    output with one TMA store per tile and reads such an input through a single tensor map (csrc/tmaps.cuh)."""
    return torch.empty((B, Cc, (T + 3) // 4 * 4), dtype=torch.float32, device=device)[:, :, :T]


def snake_conv3(w: PackedConv3, x: torch.Tensor, sigmoid: bool = False, pre_activated: bool = False,
                post_alpha: Optional[torch.Tensor] = None, padded_out: bool = False) -> torch.Tensor:
    """One block of the importance subnet, x [B,Cin,T] -> [B,Cout,T]: vrvq_snake_conv3_tc_f32 (tcgen05 3xTF32 implicit GEMM) for
    the wide blocks, vrvq_snake_conv3_f32 (CUDA cores) otherwise; VRVQ_SUBNET_IMPL=cuda forces the latter.
    `pre_activated` / `post_alpha` (tensor-core block only): the input already went through Snake / store the output through the
    next block's Snake (see importance_subnet).  `padded_out`: the result is a view of a row-padded buffer (_padded_rows)."""
    require_cuda_f32(x, "x")
    if x.dim() != 3 or x.shape[1] != w.cin:
        raise VrvqError(f"x must be [B, {w.cin}, T], got {tuple(x.shape)}")
    if x.device != w.device:
        raise VrvqError(f"x is on {x.device} but the packed weights are on {w.device}")
    B, _, T = x.shape
    y = _padded_rows(B, w.cout, T, x.device) if padded_out else torch.empty((B, w.cout, T), dtype=torch.float32, device=x.device)
    if B * T == 0:
        return y
    _check_view(x, "x")
    if not sigmoid and _tc_block_ok(w, x):
        with torch.cuda.device(x.device):
            check(_lib.lib().vrvq_snake_conv3_tc_f32(x.data_ptr(), x.stride(0), x.stride(1), None if pre_activated else w.alpha.data_ptr(),
                                                     w.packed_tc.data_ptr(), w.bias.data_ptr(), ptr(post_alpha), B, w.cin, w.cout, T, y.data_ptr(),
                                                     y.stride(0), y.stride(1), current_stream_ptr(x.device)), "vrvq_snake_conv3_tc_f32")
        _lib.count_launch()
        return y
    if pre_activated or post_alpha is not None:
        raise VrvqError("pre_activated / post_alpha are options of the tensor-core block")
    with torch.cuda.device(x.device):
        check(_lib.lib().vrvq_snake_conv3_f32(x.data_ptr(), x.stride(0), x.stride(1), w.alpha.data_ptr(), w.packed.data_ptr(),
                                              w.bias.data_ptr(), B, w.cin, w.cout, T, int(bool(sigmoid)), y.data_ptr(), y.stride(0),
                                              y.stride(1), current_stream_ptr(x.device)), "vrvq_snake_conv3_f32")
    _lib.count_launch()
    return y


def snake(x: torch.Tensor, alpha: torch.Tensor, padded_out: bool = False) -> torch.Tensor:
    """vrvq_snake_f32: x + sin(alpha x)^2 / (alpha + 1e-9) per channel (models/layers.py:25-31)."""
    require_cuda_f32(x, "x")
    _check_view(x, "x")
    B, Cc, T = x.shape
    y = _padded_rows(B, Cc, T, x.device) if padded_out else torch.empty_like(x, memory_format=torch.contiguous_format)
    if y.numel():
        with torch.cuda.device(x.device):
            check(_lib.lib().vrvq_snake_f32(x.data_ptr(), x.stride(0), x.stride(1), alpha.data_ptr(), B, Cc, T, y.data_ptr(), y.stride(0), y.stride(1),
                                            current_stream_ptr(x.device)), "vrvq_snake_f32")
        _lib.count_launch()
    return y


def subnet_tail(blocks, x: torch.Tensor, pre_activated: bool = False) -> torch.Tensor:
    """vrvq_subnet_tail_f32: the last three blocks (128 -> 32 -> 8 -> 1) and the sigmoid in one launch, x [B,128,T] -> [B,1,T]."""
    require_cuda_f32(x, "x")
    b0, b1, b2 = blocks
    if x.dim() != 3 or x.shape[1] != b0.cin:
        raise VrvqError(f"x must be [B, {b0.cin}, T], got {tuple(x.shape)}")
    B, _, T = x.shape
    y = torch.empty((B, 1, T), dtype=torch.float32, device=x.device)
    if B * T == 0:
        return y
    _check_view(x, "x")
    with torch.cuda.device(x.device):
        check(_lib.lib().vrvq_subnet_tail_f32(x.data_ptr(), x.stride(0), x.stride(1), int(bool(pre_activated)), b0.cin, b1.cin, b2.cin, b0.alpha.data_ptr(),
                                              b0.packed.data_ptr(), b0.bias.data_ptr(), b1.alpha.data_ptr(), b1.packed.data_ptr(), b1.bias.data_ptr(),
                                              b2.alpha.data_ptr(), b2.packed.data_ptr(), b2.bias.data_ptr(), B, T, y.data_ptr(), y.stride(0),
                                              current_stream_ptr(x.device)), "vrvq_subnet_tail_f32")
    _lib.count_launch()
    return y


def _tail_ok(blocks) -> bool:
    import os

    if len(blocks) < 3 or os.environ.get("VRVQ_SUBNET_IMPL", "") == "cuda":
        return False
    b0, b1, b2 = blocks[-3:]
    return b2.cout == 1 and b0.cout == b1.cin and b1.cout == b2.cin and bool(_lib.lib().vrvq_subnet_tail_usable(b0.cin, b1.cin, b2.cin))


def importance_subnet(blocks, x: torch.Tensor) -> torch.Tensor:
    """models/importance_subnet.py:38-44 on packed blocks: tensor-core launches for the wide blocks (csrc/subnet_tc.cu), one launch for
    the narrow tail 128 -> 32 -> 8 -> 1 with the sigmoid (vrvq_subnet_tail_f32); other shapes chain vrvq_snake_conv3_f32.
    A run of tensor-core blocks evaluates each Snake once: the first one's input goes through vrvq_snake_f32, every later block gets its
    activation from its predecessor's epilogue (`post_alpha`) -- inside a block the activation would be recomputed for each of the
    Cout / 128 output tiles (8 times for 1024 -> 1024: 250 of 970 us at config-2 size)."""
    blocks = list(blocks)
    n_head = len(blocks) - 3 if _tail_ok(blocks) else len(blocks)
    pre = False
    for i in range(n_head):
        w = blocks[i]
        last = i == len(blocks) - 1
        tc = (not last) and _tc_block_ok(w, x)
        if tc and not pre:
            x = snake(x, w.alpha, padded_out=True)
            pre = True
        nxt = blocks[i + 1] if i + 1 < len(blocks) else None
        post = None
        if tc and nxt is not None and ((i + 1 < n_head and i + 1 < len(blocks) - 1 and nxt.packed_tc is not None) or i + 1 == n_head):
            post = nxt.alpha  # the consumer takes activated input: a tensor-core block (its batch pitch Cout * padded T is a multiple of 4) or the tail
        if tc:
            y = snake_conv3(w, x, pre_activated=True, post_alpha=post, padded_out=True)  # (intermediate activations: rows padded to 16 bytes)
            if post is not None and i + 1 < n_head and not _tc_block_ok(nxt, y):  # cannot happen for outputs of these widths; keep the chain correct anyway
                raise VrvqError("internal: post-activated output handed to a block that cannot take it")
            x, pre = y, post is not None
        else:
            x = snake_conv3(w, x, sigmoid=last)
            pre = False
    if n_head < len(blocks):
        x = subnet_tail(blocks[n_head:], x, pre_activated=pre)
    return x
