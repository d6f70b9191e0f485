"""Building blocks UPSTREAM of the fused RVQ kernel: the DAC encoder conv stack (plain PyTorch/cuDNN on purpose, SURVEY.md
section 2 rows 5, 6: the accelerated path starts at the latent) and the importance subnet, whose eval forward runs on the
fused Snake + k=3 conv kernels of csrc/subnet.cu (`forward_torch` keeps the differentiable
PyTorch formulation for training).  The parameter layout is the reference's, so a reference checkpoint loads unchanged:
  weight-normed convs expose `weight_g`, `weight_v`, `bias`   (models/layers.py:17-18, old-style weight_norm)
  Snake1d exposes `alpha` [1,C,1]                               (models/layers.py:35-41)
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class WNConv1d(nn.Module):
    """Conv1d with w = v * g/||v|| (per output channel).  Same state-dict keys as
    torch.nn.utils.weight_norm(nn.Conv1d(...)) in the reference (models/layers.py:17-18)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.stride, self.padding, self.dilation = stride, padding, dilation
        v = torch.empty(out_channels, in_channels, kernel_size)
        nn.init.kaiming_uniform_(v, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_channels * kernel_size)
        # registration order bias, g, v matches the reference's state-dict order
        self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).view(-1, 1, 1))
        self.weight_v = nn.Parameter(v)

    def effective_weight(self):
        return torch._weight_norm(self.weight_v, self.weight_g, 0)

    def forward(self, x):
        return F.conv1d(x, self.effective_weight(), self.bias, self.stride, self.padding, self.dilation)


class Snake1d(nn.Module):
    """x + sin^2(alpha x) / alpha  (models/layers.py:25-41)."""

    def __init__(self, channels):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones(1, channels, 1))

    def forward(self, x):
        a = self.alpha
        return x + (a + 1e-9).reciprocal() * torch.sin(a * x).pow(2)


class ResidualUnit(nn.Module):
    def __init__(self, dim=16, dilation=1):
        super().__init__()
        pad = ((7 - 1) * dilation) // 2
        self.block = nn.Sequential(
            Snake1d(dim), WNConv1d(dim, dim, kernel_size=7, dilation=dilation, padding=pad),
            Snake1d(dim), WNConv1d(dim, dim, kernel_size=1))

    def forward(self, x):
        y = self.block(x)
        trim = (x.shape[-1] - y.shape[-1]) // 2
        if trim > 0:
            x = x[..., trim:-trim]
        return x + y


class EncoderBlock(nn.Module):
    def __init__(self, dim=16, stride=1):
        super().__init__()
        self.block = nn.Sequential(
            ResidualUnit(dim // 2, dilation=1), ResidualUnit(dim // 2, dilation=3), ResidualUnit(dim // 2, dilation=9),
            Snake1d(dim // 2),
            WNConv1d(dim // 2, dim, kernel_size=2 * stride, stride=stride, padding=math.ceil(stride / 2)))

    def forward(self, x):
        return self.block(x)


class Encoder(nn.Module):
    """DAC encoder (models/dac_vrvq.py:19-48): strides [2,4,8,8] -> hop 512, latent 1024.
    `return_feat` taps the output of the last EncoderBlock for the importance subnet (:39-48)."""

    def __init__(self, d_model=64, strides=(2, 4, 8, 8), latent_dim=512):
        super().__init__()
        layers = [WNConv1d(1, d_model, kernel_size=7, padding=3)]
        for s in strides:
            d_model *= 2
            layers.append(EncoderBlock(d_model, stride=s))
        layers += [Snake1d(d_model), WNConv1d(d_model, latent_dim, kernel_size=3, padding=1)]
        self.block = nn.Sequential(*layers)

    def forward(self, x, return_feat=False):
        feat = None
        tap = len(self.block) - 3
        for i, layer in enumerate(self.block):
            x = layer(x)
            if i == tap:
                feat = x
        return (x, feat) if return_feat else x


class ImportanceSubnet(nn.Module):
    """Importance map producer (models/importance_subnet.py:6-45): 6 x (Snake, k=3 conv) + sigmoid -> [B,1,T].

    Eval forward on a CUDA fp32 tensor runs this package's kernels (no PyTorch op in between, five
    launches for the shipped 1024-wide net): a Snake pre-pass, the three wide blocks as tcgen05 3xTF32 implicit GEMMs
    and the tail 128 -> 32 -> 8 -> 1 + sigmoid in one launch (all csrc/subnet_tc.cu); the generic fp32 block of csrc/subnet.cu
    serves every other shape.  There is no fallback on that path -- a missing libvrvq.so raises.
    The differentiable PyTorch formulation (`forward_torch`) serves training, which is outside this package's scope."""

    def __init__(self, d_input, d_feat, intermediate_channels=(512, 128, 32, 8), out_channels=1, detach_input=False):
        super().__init__()
        self.in_block = nn.Sequential(Snake1d(d_input), WNConv1d(d_input, d_feat, kernel_size=3, padding=1))
        cin = [d_feat] + list(intermediate_channels)
        cout = list(intermediate_channels) + [out_channels]
        self.blocks = nn.ModuleList(
            [nn.Sequential(Snake1d(i), WNConv1d(i, o, kernel_size=3, padding=1)) for i, o in zip(cin, cout)])
        self.detach_input = detach_input
        self._packed = None
        self._packed_key = None
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module.invalidate_packed())

    def _all_blocks(self):
        return [self.in_block] + list(self.blocks)

    def packed_blocks(self, device):
        """Fold weight-norm on the CPU (as the reference's hook does, models/layers.py:17-18) and pack once per
        parameter version."""
        from . import ops

        # every parameter's storage address + in-place version; `.data` surgery is invisible to it: call invalidate_packed()
        key = (tuple((p.data_ptr(), p._version) for p in self.parameters()), str(device))
        if self._packed is None or self._packed_key != key:
            self._packed = [ops.PackedConv3(snake.alpha, ops.fold_weight_norm(conv.weight_v, conv.weight_g), conv.bias, device)
                            for snake, conv in self._all_blocks()]
            self._packed_key = key
        return self._packed

    def invalidate_packed(self):
        self._packed, self._packed_key = None, None

    def _apply(self, fn, recurse=True):
        r = super()._apply(fn, recurse)
        self.invalidate_packed()
        return r

    def forward_torch(self, x):
        if self.detach_input:
            x = x.detach()
        x = self.in_block(x)
        for blk in self.blocks:
            x = blk(x)
        return torch.sigmoid(x)

    def forward(self, x):
        if self.training:
            return self.forward_torch(x)
        from . import _lib, ops

        _lib.require_cuda_f32(x, "feat_enc")
        return ops.importance_subnet(self.packed_blocks(x.device), x.detach())
