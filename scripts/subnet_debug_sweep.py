"""Profiling sweep of the tensor-core subnet block (csrc/subnet_tc.cu) under its VRVQ_SUBNET_DEBUG knobs: which role's chain bounds a tile.
python scripts/subnet_debug_sweep.py [modes ...]   (config-2 size, blocks 1024->1024 and 1024->512, pre-activated input)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vrvq_b200 import ops

def timeit(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)])) * 1e3

B, T = 16, 862
rng = np.random.Generator(np.random.PCG64(1))
x0 = torch.from_numpy(rng.normal(size=(B, 1024, T)).astype(np.float32)).cuda()
x = ops._padded_rows(B, 1024, T, "cuda")  # as inside ops.importance_subnet: rows padded to 16 bytes, outputs through the TMA store
x.copy_(x0)
post = [torch.full((1024,), 0.9, device="cuda"), None]  # block 0 stores through block 1's Snake
blocks = []
for cout in (1024, 512):
    w = torch.from_numpy((rng.normal(size=(cout, 1024, 3)) / 55).astype(np.float32))
    blocks.append(ops.PackedConv3(torch.ones(1024), w, torch.zeros(cout), "cuda"))
modes = [int(a) for a in sys.argv[1:]] or [0]
for mode in modes:
    os.environ["VRVQ_SUBNET_DEBUG"] = str(mode)
    us = [timeit(lambda: ops.snake_conv3(b, x, pre_activated=True, post_alpha=pa, padded_out=True)) for b, pa in zip(blocks, post)]
    print(f"debug {mode:3d}: 1024->1024 {us[0]:8.1f} us   1024->512 {us[1]:8.1f} us", flush=True)
    if os.environ.get("SWEEP_TRACE"):  # per-role cycle counters of block 0 (stderr), first block shape only
        os.environ["VRVQ_SUBNET_TRACE"] = "1"
        ops.snake_conv3(blocks[0], x, pre_activated=True, post_alpha=post[0], padded_out=True)
        torch.cuda.synchronize()
        del os.environ["VRVQ_SUBNET_TRACE"]
