"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (vrvq_b200/).

Imports the *unmodified* reference quantizer from a checkout of lixinghe1999/VRVQ
(default /root/reference, override with VRVQ_REFERENCE_ROOT) so that fixtures can be
generated from it and the C restatement in oracle/rvq_oracle.c can be pinned against it.

The reference imports `audiotools` and `torchmetrics` at module scope
(models/utils.py:6-7, models/dac_vrvq.py:11) although nothing on the RVQ path uses
them, and neither package is installed in this image.  We register three empty stand-in
modules before importing; the reference sources themselves are executed as they lie.

The reference tree does not exist on the GPU box: callers must check `available()`.
"""
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("VRVQ_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "quantize.py"))


def load():
    """Return a namespace with the reference classes/functions on the RVQ path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch

    for name in ("audiotools", "audiotools.ml", "torchmetrics"):
        sys.modules.setdefault(name, types.ModuleType(name))

    class _BaseModel(torch.nn.Module):  # stands in for audiotools.ml.BaseModel
        @property
        def device(self):
            return next(self.parameters()).device

    sys.modules["audiotools.ml"].BaseModel = _BaseModel
    sys.modules["audiotools"].ml = sys.modules["audiotools.ml"]
    sys.modules["audiotools"].AudioSignal = object
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    warnings.filterwarnings("ignore", category=FutureWarning)
    warnings.filterwarnings("ignore", category=UserWarning)
    from models import quantize as q  # noqa: E402
    from models import utils as u  # noqa: E402
    from models import dac_vrvq as d  # noqa: E402
    from models import dac_base as db  # noqa: E402

    ns = types.SimpleNamespace(
        VectorQuantize=q.VectorQuantize,
        ResidualVectorQuantize=q.ResidualVectorQuantize,
        VBRResidualVectorQuantize=q.VBRResidualVectorQuantize,
        generate_mask_hard=u.generate_mask_hard,
        generate_mask_ste=u.generate_mask_ste,
        cal_bpf_from_mask=u.cal_bpf_from_mask,
        DAC_VRVQ=d.DAC_VRVQ,
        Encoder=d.Encoder,
        DACFile=db.DACFile,
    )
    return ns
