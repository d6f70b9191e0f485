"""Drop-in mirrors of the reference quantizer classes (models/quantize.py) on top of the fused sm_100a kernel.

Same class names, constructor kwargs, forward signatures, return dict keys and state-dict keys
(`quantizers.{i}.in_proj.weight_g/_v/bias`, `.out_proj.*`, `.codebook.weight`, `imp_subnet.*`), so a
reference checkpoint loads unchanged and `DAC_VRVQ.encode(n_quantizers, level)` keeps working.

Scope: inference (`.eval()`), CUDA, float32.  Training-mode forward (quantizer dropout, random levels,
straight-through gradients: quantize.py:175-180, 374-386, 405-414) is outside the accelerated path and
raises NotImplementedError; CPU tensors raise VrvqError.  There is no PyTorch fallback.
"""
from typing import Union

import torch
import torch.nn as nn

from . import ops
from ._lib import CD, VrvqError
from .layers import ImportanceSubnet, WNConv1d


def _param_key(module: nn.Module):
    """Cache key of a module's packed weights: storage address and in-place version counter of EVERY parameter.
    What it sees: optimizer steps / copy_ / load_state_dict (bump `_version`), `.to()` / `.data = new` (move the storage).
    What it cannot see: writes through `.data` or under `torch.no_grad` views that bypass the counter (`p.data.mul_(2)`) --
    after such surgery call `invalidate_packed()` (load_state_dict and `.to()/.cuda()/.float()` call it themselves)."""
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


class _PackedCacheMixin:
    """Packed-weight cache shared by the quantizer mirrors: explicit invalidation + hooks on the nn.Module paths that
    rewrite parameters wholesale."""

    def _init_packed_cache(self):
        self._packed = None
        self._packed_key = None
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module.invalidate_packed())

    def invalidate_packed(self):
        """Drop the device-side weight blob; the next forward folds and packs the current parameters again."""
        self._packed = None
        self._packed_key = None
        for child in self.children():
            for m in child.modules():
                if isinstance(m, _PackedCacheMixin) and m is not self:
                    m._packed, m._packed_key = None, None

    def _apply(self, fn, recurse=True):
        r = super()._apply(fn, recurse)
        self.invalidate_packed()
        return r


class VectorQuantize(_PackedCacheMixin, nn.Module):
    """One RVQ stage (models/quantize.py:21-103).  forward() runs the fused kernel with a single stage."""

    def __init__(self, input_dim: int, codebook_size: int, codebook_dim: int):
        super().__init__()
        self.codebook_size = codebook_size
        self.codebook_dim = codebook_dim
        self.in_proj = WNConv1d(input_dim, codebook_dim, kernel_size=1)
        self.out_proj = WNConv1d(codebook_dim, input_dim, kernel_size=1)
        self.codebook = nn.Embedding(codebook_size, codebook_dim)
        self._init_packed_cache()

    def folded(self):
        """(w_in [8,D], b_in [8], w_out [D,8], b_out [D], codebook [K,8]) on the CPU, fp32."""
        return (ops.fold_weight_norm(self.in_proj.weight_v, self.in_proj.weight_g)[:, :, 0],
                self.in_proj.bias.detach().to("cpu", torch.float32),
                ops.fold_weight_norm(self.out_proj.weight_v, self.out_proj.weight_g)[:, :, 0],
                self.out_proj.bias.detach().to("cpu", torch.float32),
                self.codebook.weight.detach().to("cpu", torch.float32))

    def _weights(self, device):
        key = (_param_key(self), str(device))
        if self._packed is None or self._packed_key != key:
            parts = [t.unsqueeze(0) for t in self.folded()]
            self._packed = ops.PackedWeights(*parts, device=device)
            self._packed_key = key
        return self._packed

    def forward(self, z, loss_per_frame=False):
        """-> (z_q [B,D,T], commitment_loss, codebook_loss, indices [B,T], z_e [B,8,T])   (quantize.py:42-79)"""
        if self.training:
            raise NotImplementedError("vrvq_b200 accelerates the eval forward only; call .eval()")
        w = self._weights(z.device)
        out = ops.rvq_encode(w, z, n_run=1, want_loss_pf=True)
        loss = out.loss_pf[:, 0, :]  # [B,T] = mean over the 8 channels
        if not loss_per_frame:
            loss = loss.mean(dim=1)  # quantize.py:69: mean over [1, 2]
        return out.z_q, loss, loss.clone(), out.codes[:, 0, :], out.latents

    def embed_code(self, embed_id):
        return torch.nn.functional.embedding(embed_id, self.codebook.weight)

    def decode_code(self, embed_id):
        return self.embed_code(embed_id).transpose(1, 2)


class ResidualVectorQuantize(_PackedCacheMixin, nn.Module):
    """models/quantize.py:106-285."""

    def __init__(self, input_dim: int = 512, n_codebooks: int = 9, codebook_size: int = 1024,
                 codebook_dim: Union[int, list] = 8, quantizer_dropout: float = 0.0):
        super().__init__()
        if isinstance(codebook_dim, int):
            codebook_dim = [codebook_dim for _ in range(n_codebooks)]
        if any(c != CD for c in codebook_dim):
            raise VrvqError(f"vrvq_b200 supports codebook_dim={CD} only (all reference configs, conf/base.yml:11)")
        self.n_codebooks = n_codebooks
        self.codebook_dim = codebook_dim
        self.codebook_size = codebook_size
        self.input_dim = input_dim
        self.quantizers = nn.ModuleList([VectorQuantize(input_dim, codebook_size, codebook_dim[i]) for i in range(n_codebooks)])
        self.quantizer_dropout = quantizer_dropout
        self._init_packed_cache()

    # ---- weights ----
    def packed_weights(self, device) -> ops.PackedWeights:
        """Fold weight-norm on the CPU (bit-identical to the reference's hook) and pack once per parameter version
        (key: `_param_key`, every stage parameter's storage address and version counter; see there for what it cannot
        detect and `invalidate_packed()`)."""
        key = (tuple((p.data_ptr(), p._version) for q in self.quantizers for p in q.parameters()), device)
        if self._packed is None or self._packed_key != key:
            cols = list(zip(*[q.folded() for q in self.quantizers]))
            self._packed = ops.PackedWeights(*[torch.stack(c) for c in cols], device=device)
            self._packed_key = key
        return self._packed

    def _check_input(self, z):
        if self.training:
            raise NotImplementedError(
                "vrvq_b200 accelerates the eval forward only (training-time quantizer dropout, quantize.py:175-180, is out of scope); call .eval()")
        if z.dim() != 3 or z.shape[1] != self.input_dim:
            raise VrvqError(f"z must be [B, {self.input_dim}, T], got {tuple(z.shape)}")

    # ---- forward (quantize.py:136-214) ----
    def forward(self, z, n_quantizers: int = None):
        self._check_input(z)
        n = self.n_codebooks if n_quantizers is None else min(int(n_quantizers), self.n_codebooks)
        if n < 1:
            raise RuntimeError("stack expects a non-empty TensorList")  # what torch.stack([]) raises at quantize.py:204
        w = self.packed_weights(z.device)
        out = ops.rvq_encode(w, z, n_run=n)
        loss = (out.loss_sum[0] / max(out.frames, 1)).to(torch.float32)
        return {"z_q": out.z_q, "codes": out.codes, "latents": out.latents, "commitment_loss": loss, "codebook_loss": loss.clone()}

    # ---- decode side (quantize.py:217-249) ----
    def from_codes(self, codes: torch.Tensor, return_z_q_is=False):
        w = self.packed_weights(codes.device)
        z_q, z_p, z_q_is = ops.from_codes(w, codes, want_z_q_is=return_z_q_is)
        if return_z_q_is:
            return z_q, z_p, codes, z_q_is
        return z_q, z_p, codes

    # ---- quantize.py:251-285: re-quantise stored latents ----
    def from_latents(self, latents: torch.Tensor):
        """latents [B, 8n, T] -> (z_q [B,D,T], z_p [B,8n,T], codes [B,n,T]): per-stage nearest-neighbour search on the
        given projected latents, then gather + out_proj + sum."""
        w = self.packed_weights(latents.device)
        codes = ops.search_latents(w, latents)
        z_q, z_p, _ = ops.from_codes(w, codes)
        return z_q, z_p, codes


class VBRResidualVectorQuantize(ResidualVectorQuantize):
    """models/quantize.py:288-449."""

    def __init__(self, *, input_dim: int = 512, n_codebooks: int = 9, codebook_size: int = 1024,
                 codebook_dim: Union[int, list] = 8, quantizer_dropout: float = 0.0, full_codebook_rate: float = 0.5,
                 level_min: float, level_max: float, level_dist: str = "uniform", detach_imp_map_input: bool = False,
                 imp2mask_alpha: float = 1.0):
        super().__init__(input_dim=input_dim, n_codebooks=n_codebooks, codebook_size=codebook_size, codebook_dim=codebook_dim,
                         quantizer_dropout=quantizer_dropout)
        self.full_codebook_rate = full_codebook_rate
        self.level_min = level_min
        self.level_max = level_max
        self.level_dist = level_dist
        self.detach_imp_map_input = detach_imp_map_input
        self.imp2mask_alpha = imp2mask_alpha
        self.imp_subnet = ImportanceSubnet(d_input=input_dim, d_feat=input_dim, intermediate_channels=[512, 128, 32, 8],
                                           out_channels=1, detach_input=detach_imp_map_input)
        # Extension: set False to skip materialising z_q_is [B,Nq,D,T] (4096*Nq bytes/frame) when the caller
        # takes `level` at encode time instead of re-masking afterwards.  Default keeps the reference's dict.
        self.return_z_q_is = True
        # Extension: allocate z_q_is with a 128-byte aligned row pitch and return a [..., :T] view (same values, not
        # contiguous).  About 14 % faster when T is not a multiple of 4 (e.g. T=862).  Default keeps the reference layout.
        self.pad_z_q_is_rows = False

    def forward(self, z: torch.Tensor, n_quantizers: int = None, feat_enc: torch.Tensor = None, level: float = None,
                imp_map: torch.Tensor = None):
        """quantize.py:328-443 (eval).  `imp_map` (extension) bypasses the subnet with a precomputed map."""
        self._check_input(z)
        Nq = self.n_codebooks
        w = self.packed_weights(z.device)
        B, D, T = z.shape
        if n_quantizers is None:  # ---- VBR mode
            assert level is not None, "level must be specified in VBR mode"
            if imp_map is None:
                imp_map = self.imp_subnet(feat_enc)  # csrc/subnet_tc.cu: Snake pre-pass, three tcgen05 blocks, fused tail (quantize.py:372)
            imp_in, lvl = imp_map.contiguous(), level
            if isinstance(level, torch.Tensor):
                # the reference broadcasts `imp_map [B,1,T] * level` (quantize.py:389): a scalar tensor or [B,1,1] is a
                # per-item level (the kernel reads it); every other shape goes through the same torch broadcast, so e.g.
                # [1,1,T] scales per frame and a shape the reference rejects is rejected here with the same error
                lv = level.to(device=z.device, dtype=torch.float32)
                if lv.numel() == 1 or tuple(lv.shape) == (B, 1, 1):
                    lvl = lv.reshape(-1)
                else:  # pre-multiply (the same fp32 product the reference forms first), then level = 1 is exact
                    scaled = imp_map * lv
                    if scaled.shape != imp_map.shape:
                        raise RuntimeError(f"level of shape {tuple(lv.shape)} does not broadcast against imp_map {tuple(imp_map.shape)} "
                                           f"without changing its shape (quantize.py:389-395 would fail in generate_mask)")
                    imp_in, lvl = scaled.contiguous(), 1.0
            out = ops.EncodeOutputs(B, D, T, Nq, z.device, z_q=True, z_q_is=self.return_z_q_is, latents=True, mask=True,
                                    pad_z_q_is_rows=self.pad_z_q_is_rows)
            if B * T:
                ops.rvq_encode_into(w, z, out, Nq, imp_in, lvl, zero_accum=False)
            mask_imp, imp_out, z_q, z_q_is = out.mask, imp_map, out.z_q, out.z_q_is
        else:  # ---- CBR mode inside the VBR model (quantize.py:397-400)
            n = min(int(n_quantizers), Nq)
            if n < 1:
                raise RuntimeError("stack expects a non-empty TensorList")
            if 1 < n < Nq:
                # the reference fails here: [B,n,D,T] * [B,Nq,1,T] does not broadcast (quantize.py:400,421)
                raise RuntimeError(f"The size of tensor a ({n}) must match the size of tensor b ({Nq}) at non-singleton dimension 1")
            out = ops.EncodeOutputs(B, D, T, n, z.device, z_q=True, z_q_is=self.return_z_q_is, latents=True, mask=False)
            if B * T:
                ops.rvq_encode_into(w, z, out, n, None, None, zero_accum=False)
            z_q, z_q_is = out.z_q, out.z_q_is
            if n == 1 and Nq > 1:
                # reference quirk kept for parity: the single z_q_0 broadcasts against the all-ones [B,Nq,T] mask
                # and is summed Nq times (quantize.py:400,420-421).  The loss buffers are zero beyond stage 0, so
                # the losses are not scaled.
                z_q = z_q * float(Nq)
            mask_imp = torch.ones((B, Nq, T), dtype=torch.float32, device=z.device)
            imp_out = None
        loss = (out.loss_sum[0] / max(out.frames, 1)).to(torch.float32)
        return {"z_q": z_q, "z_q_is": z_q_is, "codes": out.codes, "latents": out.latents, "commitment_loss": loss,
                "codebook_loss": loss.clone(), "imp_map": imp_out, "mask_imp": mask_imp,
                # extension (not in the reference dict): exact per-stage kept-frame counts for bits-per-frame
                "kept_frames": out.kept}

    def from_codes(self, codes: torch.Tensor, return_z_q_is=False):
        raise NotImplementedError  # quantize.py:445-446

    def from_latents(self, latents: torch.Tensor):
        raise NotImplementedError  # quantize.py:448-449

    def from_codes_masked(self, codes: torch.Tensor, mask: torch.Tensor):
        """Extension: VBR-aware decode the reference left as a TODO (quantize.py:228): z_q = sum_k mask_k * z_q_k."""
        w = self.packed_weights(codes.device)
        z_q, _, _ = ops.from_codes(w, codes, mask=mask, want_z_p=False)
        return z_q
