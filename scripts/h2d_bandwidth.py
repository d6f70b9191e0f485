"""Host -> device copy bandwidth of the bench's e2e input (56.5 MB of latents) from pageable-pinned vs write-combined pinned memory,
one process per GPU (run under torchrun to see the aggregate limit of the box).  python scripts/h2d_bandwidth.py [--wc-only]"""
import ctypes, os, sys, time
import torch

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
N = 16 * 1024 * 862


def wc_pinned(n_floats):
    from cuda import cudart
    err, ptr = cudart.cudaHostAlloc(n_floats * 4, cudart.cudaHostAllocWriteCombined)
    assert int(err) == 0, err
    buf = (ctypes.c_float * n_floats).from_address(int(ptr))
    return torch.frombuffer(buf, dtype=torch.float32)


def bw(bufs, steps=40):
    d = torch.empty(N, device=dev)
    for b in bufs:
        d.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        d.copy_(bufs[i % len(bufs)], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return N * 4 * steps / (e0.elapsed_time(e1) * 1e-3) / 1e9


if world > 1:
    torch.distributed.init_process_group("nccl")
src = torch.randn(N)
pinned = [src.clone().pin_memory() for _ in range(2)]
wcs = [wc_pinned(N) for _ in range(2)]
for w in wcs:
    w.copy_(src)
res = {"pinned": bw(pinned), "wc": bw(wcs), "pinned2": bw(pinned), "wc2": bw(wcs), "wc_is_pinned": wcs[0].is_pinned()}
t = torch.tensor([res["pinned"], res["wc"], res["pinned2"], res["wc2"]], device=dev)
if world > 1:
    mn = t.clone(); torch.distributed.all_reduce(mn, op=torch.distributed.ReduceOp.MIN)
    sm = t.clone(); torch.distributed.all_reduce(sm, op=torch.distributed.ReduceOp.SUM)
    if rank == 0:
        print(f"world {world}: GB/s per GPU (min over ranks) pinned {mn[0]:.1f} wc {mn[1]:.1f} pinned {mn[2]:.1f} wc {mn[3]:.1f}; aggregate pinned {sm[0]:.0f} wc {sm[1]:.0f}; wc is_pinned {res['wc_is_pinned']}")
else:
    print(f"world 1: GB/s pinned {res['pinned']:.1f} wc {res['wc']:.1f} pinned {res['pinned2']:.1f} wc {res['wc2']:.1f}; wc is_pinned {res['wc_is_pinned']}")
