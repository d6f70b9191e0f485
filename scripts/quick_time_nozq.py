"""Scratch: config-3 / config-4-shape launches with and without the final z_q GEMM (cost of the final phase)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vrvq_b200 import ops
from tests.golden import gen_inputs as gi

def run(B, T, Nq, zq, n_run=None, iters=20):
    n_run = n_run or Nq
    sd = gi.torch_state_dict(gi.make_state_dict(1, Nq, 1024))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    zs = [torch.randn(B, 1024, T, device="cuda") for _ in range(4)]
    imp = torch.rand(B, 1, T, device="cuda")
    out = ops.EncodeOutputs(B, 1024, T, Nq, "cuda", z_q=zq, z_q_is=False, latents=True, mask=True)
    for i in range(3):
        ops.rvq_encode_into(pw, zs[i % 4], out, n_run, imp if n_run == Nq else None, 0.5 if n_run == Nq else None)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(iters):
        evs[i][0].record()
        ops.rvq_encode_into(pw, zs[i % 4], out, n_run, imp if n_run == Nq else None, 0.5 if n_run == Nq else None, zero_accum=False)
        evs[i][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    print(f"B={B} T={T} Nq={Nq} n_run={n_run} z_q={zq}: median {ts[len(ts)//2]*1e3:.1f} us", flush=True)

run(64, 862, 28, True)
run(64, 862, 28, False)
run(64, 862, 28, False, n_run=8)
run(64, 862, 28, False, n_run=16)
run(32, 5168, 8, True)
run(32, 5168, 8, False)
