"""Host-side cost per call of the Python mirror (no GPU sync inside the loop) vs the device time of the kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vrvq_b200
from vrvq_b200 import ops
from tests.golden import gen_inputs as gi
sd = gi.torch_state_dict(gi.make_state_dict(0, 8, 1024))
m = vrvq_b200.VBRResidualVectorQuantize(input_dim=1024, n_codebooks=8, codebook_size=1024, codebook_dim=8, level_min=0.125, level_max=6.0)
m.load_state_dict(sd, strict=False); m = m.cuda().eval()
z = torch.randn(16, 1024, 862, device="cuda"); imp = torch.rand(16, 1, 862, device="cuda")
for _ in range(3): r = m(z, level=0.5, imp_map=imp)
torch.cuda.synchronize()
n = 50
t0 = time.perf_counter()
for _ in range(n): r = m(z, level=0.5, imp_map=imp)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"module forward: host {1e6*(t1-t0)/n:.0f} us/call, wall incl. drain {1e6*(t2-t0)/n:.0f} us/call")
pw = m.packed_weights(z.device)
out = ops.EncodeOutputs(16, 1024, 862, 8, "cuda", z_q=True, z_q_is=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n): ops.rvq_encode_into(pw, z, out, 8, imp, 0.5, zero_accum=False)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"rvq_encode_into: host {1e6*(t1-t0)/n:.0f} us/call, wall incl. drain {1e6*(t2-t0)/n:.0f} us/call")
t0 = time.perf_counter()
for _ in range(n): k = m.packed_weights(z.device)
print(f"packed_weights() cache check: {1e6*(time.perf_counter()-t0)/n:.0f} us/call")
