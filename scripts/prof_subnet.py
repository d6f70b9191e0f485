"""Minimal launch sequence for ncu: the six importance-subnet blocks at config-2 size (B=16, T=862), twice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden import gen_inputs as gi
from vrvq_b200.layers import ImportanceSubnet

m = ImportanceSubnet(d_input=1024, d_feat=1024)
m.load_state_dict(gi.torch_state_dict(gi.make_subnet_state_dict(31, 1024, 1024)), strict=True)
m = m.cuda().eval()
x = torch.from_numpy(gi.make_latents(5, 16, 1024, 862, 1.0)).cuda()
for _ in range(2):
    y = m(x)
torch.cuda.synchronize()
print("ok", float(y.mean()))
