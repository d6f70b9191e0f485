"""vrvq_b200 -- the residual-vector-quantisation hot path of lixinghe1999/VRVQ as one fused sm_100a kernel,
behind the reference's own Python API.  See DESIGN.md; the C ABI is include/vrvq.h.
"""
from .quantize import ResidualVectorQuantize, VBRResidualVectorQuantize, VectorQuantize  # noqa: F401
from .utils import cal_bpf_from_mask, generate_mask_hard, generate_mask_ste  # noqa: F401
from .dac_vrvq import DAC_VRVQ  # noqa: F401
from ._lib import VrvqError  # noqa: F401
from .wire import DACFile, pack_codes, unpack_codes  # noqa: F401

__all__ = ["VectorQuantize", "ResidualVectorQuantize", "VBRResidualVectorQuantize", "DAC_VRVQ", "generate_mask_hard",
           "generate_mask_ste", "cal_bpf_from_mask", "VrvqError", "DACFile", "pack_codes", "unpack_codes"]
