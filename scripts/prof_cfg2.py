"""Minimal launch sequence for ncu: a few fused encodes of the cfg2 workload (B=16,T=862,Nq=8, full dict)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vrvq_b200 import ops
from tests.golden import gen_inputs as gi

zqis = "--no-zqis" not in sys.argv
sd = gi.torch_state_dict(gi.make_state_dict(0, 8, 1024))
pw = ops.PackedWeights.from_state_dict(sd, "cuda")
B, T = 16, int(os.environ.get("PROF_T", "862"))
zs = [torch.randn(B, 1024, T, device="cuda") for _ in range(2)]
imp = torch.rand(B, 1, T, device="cuda")
out = ops.EncodeOutputs(B, 1024, T, 8, "cuda", z_q=True, z_q_is=zqis, latents=True, mask=True)
for i in range(6):
    ops.rvq_encode_into(pw, zs[i % 2], out, 8, imp, [0.25, 0.5, 1.0][i % 3], zero_accum=False)
torch.cuda.synchronize()
print("ok", out.codes[0, :, 0].tolist())
