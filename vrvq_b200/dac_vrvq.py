"""DAC_VRVQ encode-side mirror (models/dac_vrvq.py:83-213): PyTorch conv encoder (+ importance subnet) feeding the
fused RVQ kernel.  `encode(audio_data, n_quantizers, level)` keeps the reference signature and return dict.

The decoder (models/dac_vrvq.py:51-80,215-220) is not on the encode hot path and is not built here
(SURVEY.md section 2 row 6); `decode`/`forward` raise.  Reference checkpoints load with
`load_reference_state_dict`, which ignores `decoder.*` keys.
"""
import math
from typing import List, Union

import numpy as np
import torch
import torch.nn as nn

from .layers import Encoder
from .quantize import ResidualVectorQuantize, VBRResidualVectorQuantize


class DAC_VRVQ(nn.Module):
    def __init__(self, encoder_dim: int = 64, encoder_rates: List[int] = [2, 4, 8, 8], latent_dim: int = None,
                 decoder_dim: int = 1536, decoder_rates: List[int] = [8, 8, 4, 2], n_codebooks: int = 9,
                 codebook_size: Union[int, list] = 1024, codebook_dim: Union[int, list] = 8, quantizer_dropout: float = 0.0,
                 sample_rate: int = 44100, model_type: str = "VBR", full_codebook_rate: float = 0.0, level_min: float = None,
                 level_max: float = None, level_dist: str = "uniform", detach_imp_map_input: bool = False,
                 imp2mask_alpha: float = 1.0):
        super().__init__()
        self.encoder_dim, self.encoder_rates = encoder_dim, encoder_rates
        self.decoder_dim, self.decoder_rates = decoder_dim, decoder_rates
        self.sample_rate = sample_rate
        if latent_dim is None:
            latent_dim = encoder_dim * (2 ** len(encoder_rates))
        self.latent_dim = latent_dim
        self.hop_length = int(np.prod(encoder_rates))
        self.encoder = Encoder(encoder_dim, encoder_rates, latent_dim)
        self.n_codebooks, self.codebook_size, self.codebook_dim = n_codebooks, codebook_size, codebook_dim
        self.model_type = model_type
        if model_type == "CBR":
            self.quantizer = ResidualVectorQuantize(input_dim=latent_dim, n_codebooks=n_codebooks, codebook_size=codebook_size,
                                                    codebook_dim=codebook_dim, quantizer_dropout=quantizer_dropout)
        elif model_type == "VBR":
            self.quantizer = VBRResidualVectorQuantize(
                input_dim=latent_dim, n_codebooks=n_codebooks, codebook_size=codebook_size, codebook_dim=codebook_dim,
                quantizer_dropout=quantizer_dropout, full_codebook_rate=full_codebook_rate, level_min=level_min, level_max=level_max,
                level_dist=level_dist, detach_imp_map_input=detach_imp_map_input, imp2mask_alpha=imp2mask_alpha)
        else:
            raise ValueError(f"Invalid RVQ model_type: {model_type}")
        for m in self.modules():  # models/layers.py:44-49 as it acts in effect: conv biases start at zero
            if hasattr(m, "weight_v") and getattr(m, "bias", None) is not None:
                nn.init.constant_(m.bias, 0)

    @property
    def device(self):
        return next(self.parameters()).device

    def load_reference_state_dict(self, state_dict):
        """Load a reference checkpoint's `state_dict` (scripts/inference.py:41-46), skipping the decoder."""
        sd = {k: v for k, v in state_dict.items() if not k.startswith("decoder.")}
        return self.load_state_dict(sd, strict=True)

    def preprocess(self, audio_data, sample_rate):
        """Right-pad to a multiple of the hop (models/dac_vrvq.py:164-173)."""
        if sample_rate is None:
            sample_rate = self.sample_rate
        assert sample_rate == self.sample_rate
        length = audio_data.shape[-1]
        right_pad = math.ceil(length / self.hop_length) * self.hop_length - length
        return nn.functional.pad(audio_data, (0, right_pad))

    def encode(self, audio_data: torch.Tensor, n_quantizers: int = None, level: int = 1):
        """models/dac_vrvq.py:176-213: audio [B,1,S] -> quantizer dict (z_q, z_q_is, codes, latents, losses, imp_map, mask_imp)."""
        z, feat = self.encoder(audio_data, return_feat=True)
        if self.model_type == "CBR":
            return self.quantizer(z=z, n_quantizers=n_quantizers)
        return self.quantizer(z=z, n_quantizers=n_quantizers, feat_enc=feat, level=level)

    def compress(self, audio_path_or_signal, win_duration: float = 1.0, verbose: bool = False, normalize_db: float = -16,
                 n_quantizers: int = None):
        """models/dac_base.py:129-169: the reference's chunked `compress` raises NotImplementedError on its first line ("TODO: Implement
        this function"; everything below it is unreachable and needs audiotools).  Same signature, same behaviour.  The working
        container path of this package is encode() -> wire.pack_codes -> wire.DACFile (vrvq_b200/wire.py)."""
        raise NotImplementedError

    def decompress(self, obj, verbose: bool = False):
        """models/dac_base.py:242-261: raises NotImplementedError in the reference as well; wire.unpack_codes + quantizer.from_codes
        (+ from_codes_masked for VBR) is the decode-side path built here."""
        raise NotImplementedError

    def decode(self, z):
        raise NotImplementedError("the DAC decoder is outside the accelerated encode path (DESIGN.md, out of scope)")

    def forward(self, *a, **k):
        raise NotImplementedError("use encode(); decode/forward need the decoder, which is outside the accelerated path")
