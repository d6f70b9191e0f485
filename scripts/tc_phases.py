"""Scratch: per-phase cycle breakdown (VRVQ_DEBUG_PHASES=1) and timing of the tensor-core encode kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vrvq_b200 import ops
from tests.golden import gen_inputs as gi

B, T, Nq, zqis = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] == "1"
sd = gi.torch_state_dict(gi.make_state_dict(1, Nq, 1024))
pw = ops.PackedWeights.from_state_dict(sd, "cuda")
z = torch.randn(B, 1024, T, device="cuda")
imp = torch.rand(B, 1, T, device="cuda")
out = ops.EncodeOutputs(B, 1024, T, Nq, "cuda", z_q=True, z_q_is=zqis, latents=True, mask=True)
os.environ.pop("VRVQ_DEBUG_PHASES", None)
for _ in range(3):
    ops.rvq_encode_into(pw, z, out, Nq, imp, 0.5)
torch.cuda.synchronize()
os.environ["VRVQ_DEBUG_PHASES"] = os.environ.get("PHASE_MODE", "1")  # "2": production instantiation, three timestamps per CTA
ops.rvq_encode_into(pw, z, out, Nq, imp, 0.5)
torch.cuda.synchronize()
