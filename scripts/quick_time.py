"""Scratch timing of the fused encode on one GPU (device-resident inputs); bench.py is the contract version."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vrvq_b200 import ops
from tests.golden import gen_inputs as gi

def run(B, T, Nq, zqis, iters=20):
    sd = gi.torch_state_dict(gi.make_state_dict(1, Nq, 1024))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    zs = [torch.randn(B, 1024, T, device="cuda") for _ in range(4)]
    imp = torch.rand(B, 1, T, device="cuda")
    out = ops.EncodeOutputs(B, 1024, T, Nq, "cuda", z_q=True, z_q_is=zqis, latents=True, mask=True)
    for i in range(3):
        ops.rvq_encode_into(pw, zs[i % 4], out, Nq, imp, 0.5)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(iters):
        evs[i][0].record()
        ops.rvq_encode_into(pw, zs[i % 4], out, Nq, imp, 0.5, zero_accum=False)
        evs[i][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    med = ts[len(ts) // 2]
    frames = B * T
    bpf = 4096 * 2 + 4 + Nq * (8 + 4 + 32) + (4096 * Nq if zqis else 0)
    print(f"B={B} T={T} Nq={Nq} zqis={zqis}: median {med*1e3:.1f} us  min {ts[0]*1e3:.1f} us  {frames/med/1e3:.2f} Mframes/s  {frames*bpf/med/1e6:.1f} GB/s algorithmic", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        B, T, Nq, zq = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] == "1"
        run(B, T, Nq, zq)
        sys.exit(0)
    run(16, 862, 8, True)
    run(16, 862, 8, False)
    run(64, 862, 28, False)
    run(32, 5168, 8, False)
    run(16, 864, 8, True)
