"""Times the importance-subnet kernels (csrc/subnet.cu) at config-2 size (B=16, T=862, 1024 -> 1024 -> 512 -> 128 -> 32 -> 8 -> 1)
with CUDA events on the launch stream, per block and for the whole chain, next to the same module through PyTorch/cuDNN
(`forward_torch`, fp32 with TF32 off and on).  Roofline: fp32 FMA peak = SMs x 128 lanes x 2 x SM clock.

python scripts/bench_subnet.py [--out gpurun_out/subnet_bench.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden import gen_inputs as gi  # noqa: E402
from vrvq_b200 import ops  # noqa: E402
from vrvq_b200.layers import ImportanceSubnet  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)])) * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--T", type=int, default=862)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, T = a.B, a.T
    m = ImportanceSubnet(d_input=1024, d_feat=1024)
    m.load_state_dict(gi.torch_state_dict(gi.make_subnet_state_dict(31, 1024, 1024)), strict=True)
    m = m.to(dev).eval()
    x = torch.from_numpy(gi.make_latents(5, B, 1024, T, 1.0)).to(dev)
    blocks = m.packed_blocks(dev)
    props = torch.cuda.get_device_properties(dev)
    try:
        import pynvml

        pynvml.nvmlInit()
        sm_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(pynvml.nvmlDeviceGetHandleByIndex(0), pynvml.NVML_CLOCK_SM))
    except Exception:
        sm_mhz = 1965  # B200 boost clock seen in every bench run of this repo (profiles/r1*_bench.json)
    peak_tf = props.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    rows, cur = [], x
    for i, w in enumerate(blocks):
        last = i == len(blocks) - 1
        us = timeit(lambda: ops.snake_conv3(w, cur, sigmoid=last))
        flops = 2.0 * 3 * w.cin * w.cout * B * T
        rows.append(dict(block=i, cin=w.cin, cout=w.cout, us=us, gflop=flops / 1e9, tflops=flops / us / 1e6,
                         frac_fp32_peak=flops / us / 1e6 / peak_tf))
        cur = ops.snake_conv3(w, cur, sigmoid=last)
    # the launches of the actual chain (ops.importance_subnet): Snake pre-pass, tensor-core blocks on activated row-padded input, fused tail
    chain_rows = []
    if blocks[0].packed_tc is not None and len(blocks) == 6:
        xs = ops.snake(x, blocks[0].alpha, padded_out=True)
        chain_rows.append(dict(launch="snake_prepass", us=timeit(lambda: ops.snake(x, blocks[0].alpha, padded_out=True))))
        cur = xs
        for i in range(3):
            pa = blocks[i + 1].alpha
            f = lambda cur=cur, i=i, pa=pa: ops.snake_conv3(blocks[i], cur, pre_activated=True, post_alpha=pa, padded_out=True)
            chain_rows.append(dict(launch=f"tc_block_{i}", cin=blocks[i].cin, cout=blocks[i].cout, us=timeit(f)))
            cur = f()
        chain_rows.append(dict(launch="fused_tail", us=timeit(lambda: ops.subnet_tail(blocks[3:], cur, pre_activated=True))))
    chain_us = timeit(lambda: m(x))
    total = sum(r["gflop"] for r in rows) * 1e9
    res = dict(B=B, T=T, frames=B * T, sm_mhz=sm_mhz, fp32_peak_tflops=peak_tf, blocks=rows, chain_launches=chain_rows, chain_us=chain_us,
               chain_tflops=total / chain_us / 1e6, chain_frac_fp32_peak=total / chain_us / 1e6 / peak_tf,
               Mframes_per_s=B * T / chain_us)
    with torch.no_grad():
        for name, tf32 in (("torch_cudnn_fp32_us", False), ("torch_cudnn_tf32_us", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            res[name] = timeit(lambda: m.forward_torch(x))
            y_t = m.forward_torch(x)
            res[name.replace("_us", "_maxdiff_vs_kernel")] = float((y_t - m(x)).abs().max())
    print(json.dumps(res, indent=1))
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
