#!/usr/bin/env python
"""BASELINE.json configs[4]: the full encode path -- DAC encoder conv stack (PyTorch/cuDNN) feeding the fused importance-subnet
and RVQ kernels through `DAC_VRVQ.encode(audio, n_quantizers=None, level)` (models/dac_vrvq.py:176-213) -- B=32 x 5 s of
44.1 kHz audio (T=431 latent frames per item), end-to-end latent frames/s on N GPUs (4 items per GPU at N=8).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/bench_cfg5_encode.py [--steps K] [--warmup W] [--tf32-encoder] [--out gpurun_out/r2_cfg5_Ngpu.json]

A step = this rank's share of the batch: pinned host audio -> device, encode, codes + mask + kept counts back to pinned host
memory.  Timed with CUDA events, barrier + synchronize on both sides, MAX over ranks; weights are seeded random-init
(tests/golden/gen_inputs.make_dac_state_dict), audio synthetic.  The encoder runs in fp32 with TF32 off (the setting the
parity tests use); --tf32-encoder lets cuDNN use TF32 for the conv stack as PyTorch does by default.  The split
encoder / subnet / RVQ is measured on rank 0 with events in a separate untimed pass.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--level", type=float, default=1.0)
    ap.add_argument("--tf32-encoder", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = bool(args.tf32_encoder)
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32_encoder)
    torch.backends.cudnn.benchmark = True

    import vrvq_b200
    from tests.golden import gen_inputs as gi
    from vrvq_b200 import _lib

    Nq = 8
    m = vrvq_b200.DAC_VRVQ(n_codebooks=Nq, model_type="VBR", level_min=0.125, level_max=6.0, imp2mask_alpha=2.0).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(gi.torch_state_dict(gi.make_dac_state_dict(41, shapes)), strict=True)
    m = m.to(dev)
    per = args.B // world
    assert per * world == args.B, "B must divide by the number of GPUs"
    samples = int(round(args.seconds * 44100))
    x_h = torch.from_numpy(gi.make_audio(500 + rank, per, samples)).pin_memory()
    T = -(-samples // 512)
    h_codes = torch.empty((per, Nq, T), dtype=torch.int64).pin_memory()
    h_mask = torch.empty((per, Nq, T), dtype=torch.float32).pin_memory()
    h_kept = torch.empty((Nq,), dtype=torch.int64).pin_memory()

    @torch.no_grad()
    def step():
        x = m.preprocess(x_h.to(dev, non_blocking=True), 44100)
        r = m.encode(x, None, args.level)
        h_codes.copy_(r["codes"], non_blocking=True)
        h_mask.copy_(r["mask_imp"], non_blocking=True)
        h_kept.copy_(r["kept_frames"], non_blocking=True)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        r = step()
    e1.record()
    barrier()
    launches = (_lib.launch_count - l0) // args.steps
    ms = e0.elapsed_time(e1) / args.steps
    assert tuple(r["codes"].shape) == (per, Nq, T)
    # untimed split pass (rank 0's own numbers)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.no_grad():
        x = m.preprocess(x_h.to(dev), 44100)
        torch.cuda.synchronize()
        ev[0].record()
        z, feat = m.encoder(x, return_feat=True)
        ev[1].record()
        imp = m.quantizer.imp_subnet(feat)
        ev[2].record()
        m.quantizer(z=z, n_quantizers=None, feat_enc=None, level=args.level, imp_map=imp)
        ev[3].record()
        torch.cuda.synchronize()
    split = {"encoder_ms": ev[0].elapsed_time(ev[1]), "subnet_ms": ev[1].elapsed_time(ev[2]), "rvq_ms": ev[2].elapsed_time(ev[3])}
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    if rank == 0:
        line = {"workload": "cfg5", "metric": "encode_latent_frames_per_sec", "value": args.B * T / (ms_max * 1e-3), "unit": "frames/s",
                "n_gpus": world, "scaling": "strong", "B": args.B, "items_per_gpu": per, "seconds_per_item": args.seconds, "T": T,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step_max_over_ranks": ms_max, "rank0_split_ms": split,
                "vrvq_launches_per_step": launches, "encoder": "PyTorch/cuDNN fp32" + (" (TF32 allowed)" if args.tf32_encoder else " (TF32 off)"),
                "api": "vrvq_b200.DAC_VRVQ.encode(audio, None, level): pinned host audio in, codes + mask + kept counts out",
                "h2d_bytes_per_step_per_gpu": x_h.numel() * 4, "d2h_bytes_per_step_per_gpu": h_codes.numel() * 8 + h_mask.numel() * 4 + 64,
                "data": "synthetic audio N(0, 0.5), seeded random-init weights"}
        print(json.dumps(line), flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(line, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
