"""Minimal launch sequence for ncu: fused encodes without z_q_is on the config-4 shard shape (B=32, T=5168, Nq=8) or, with
PROF_CFG=3, config 3 (B=64, T=862, Nq=28)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vrvq_b200 import ops
from tests.golden import gen_inputs as gi

B, T, Nq = (64, 862, 28) if os.environ.get("PROF_CFG") == "3" else (32, 5168, 8)
sd = gi.torch_state_dict(gi.make_state_dict(0, Nq, 1024))
pw = ops.PackedWeights.from_state_dict(sd, "cuda")
z = torch.randn(B, 1024, T, device="cuda")
imp = torch.rand(B, 1, T, device="cuda")
out = ops.EncodeOutputs(B, 1024, T, Nq, "cuda", z_q=True, z_q_is=False, latents=True, mask=True)
for i in range(4):
    ops.rvq_encode_into(pw, z, out, Nq, imp, 0.5, zero_accum=False)
torch.cuda.synchronize()
print("ok", out.codes[0, :, 0].tolist())
