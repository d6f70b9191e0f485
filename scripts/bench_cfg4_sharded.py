#!/usr/bin/env python
"""BASELINE.json configs[3]: the long-form batch -- 60 s x B=256 (T=5168, 1.32 M latent frames, 5.42 GB of latents) -- cut by
batch x frame into one shard per GPU with `vrvq_b200.sharding.plan_shards`, one process per GPU, no data-path collective.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/bench_cfg4_sharded.py [--steps K] [--warmup W] [--B 256] [--T 5168] [--out gpurun_out/r2_cfg4_Ngpu.json]
  (N = 1 runs the whole batch on one GPU: the strong-scaling baseline and the bpf reference.)

Every rank generates ONLY the batch items its shard touches, on its own device, from a per-item seed (so that the 1-GPU run and
every sharded run see the same latents), runs the fused encode on its segments (strided views of whole items / frame ranges,
`sharding.encode_shard`), and the only cross-rank exchange is `sharding.reduce_counts` (kept-frame counts + loss sum: the
numerator of cal_bpf_from_mask, models/utils.py:64-73) after the timed region.  Timing: CUDA events around K passes over the
shard, barrier + synchronize on both sides, MAX over ranks; "scaling": "strong" (the total work is fixed as N grows).
Outputs per frame: codes, z_q, mask (8 292 algorithmic bytes per frame).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def item_latent(b, D, T, dev):
    g = torch.Generator(device=dev).manual_seed(100000 + b)
    return torch.randn(D, T, generator=g, device=dev)


def item_imp(b, T, dev):
    g = torch.Generator(device=dev).manual_seed(200000 + b)
    return torch.rand(1, T, generator=g, device=dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--T", type=int, default=5168)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--level", type=float, default=0.7)
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

    from tests.golden import gen_inputs as gi
    from vrvq_b200 import ops, sharding

    B, T, D, Nq = args.B, args.T, 1024, 8
    sd = gi.torch_state_dict(gi.make_state_dict(81, Nq, D))
    pw = ops.PackedWeights.from_state_dict(sd, dev)
    segs = sharding.plan_shards(B, T, world)[rank]
    # local copy of the items this rank touches: item b of the job = local index b - b_lo
    if segs:
        b_lo, b_hi = segs[0].b, segs[-1].b + 1
    else:
        b_lo, b_hi = 0, 0
    nb = b_hi - b_lo
    z = torch.empty((nb, D, T), dtype=torch.float32, device=dev)
    imp = torch.empty((nb, 1, T), dtype=torch.float32, device=dev)
    for i in range(nb):
        z[i] = item_latent(b_lo + i, D, T, dev)
        imp[i] = item_imp(b_lo + i, T, dev)
    local_segs = [sharding.Segment(s.b - b_lo, s.t0, s.t1) for s in segs]
    my_frames = sum(s.frames for s in segs)
    units = sharding.merge_whole_items(local_segs, T)

    # pre-allocated outputs (codes, z_q, mask), written in place through views; accumulators zeroed per pass
    out = ops.EncodeOutputs(nb, D, T, Nq, dev, z_q=True, z_q_is=False, latents=False, mask=True)

    def views():
        vs = []
        for u in units:
            bs, ts = (slice(u[1], u[2]), slice(0, T)) if u[0] == "items" else (slice(u[1], u[1] + 1), slice(u[2], u[3]))
            v = ops.EncodeOutputs.__new__(ops.EncodeOutputs)
            v.codes, v.z_q, v.z_q_is, v.latents = out.codes[bs, :, ts], out.z_q[bs, :, ts], None, None
            v.mask, v.loss_pf, v.accum, v.n_run = out.mask[bs, :, ts], None, out.accum, Nq
            vs.append((v, z[bs, :, ts], imp.reshape(nb, T)[bs, ts]))
        return vs

    vs = views()

    def one_pass():
        out.accum.zero_()
        for v, zz, ii in vs:
            ops.rvq_encode_into(pw, zz, v, Nq, ii, args.level, zero_accum=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_pass()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_pass()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    kept, loss = out.kept.clone(), out.loss_sum.clone()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    per_rank = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        sharding.reduce_counts(kept, loss)  # the path's only cross-rank exchange
    else:
        per_rank = [t]
    per_rank_ms = [float(x.item()) for x in per_rank]
    ms_max = max(per_rank_ms)
    if rank == 0:
        total = B * T
        bytes_pf = 4 * D + 4 + 4 * D + 8 * Nq + 4 * Nq
        peak = 6544.7
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"])
        except Exception:
            pass
        kept_l = [int(x) for x in kept.tolist()]
        line = {
            "workload": "cfg4", "metric": "rvq_latent_frames_per_sec", "value": total / (ms_max * 1e-3), "unit": "frames/s",
            "n_gpus": world, "scaling": "strong", "B": B, "T": T, "D": D, "n_codebooks": Nq, "frames_total": total,
            "frames_rank0": my_frames, "launches_per_pass_rank0": len(vs), "steps": args.steps, "warmup": args.warmup,
            "ms_per_pass_max_over_ranks": ms_max, "ms_per_pass_per_rank": per_rank_ms,
            "algorithmic_bytes_per_frame": bytes_pf,
            "hbm_frac_per_gpu": (total / world) * bytes_pf / (ms_max * 1e-3) / 1e9 / peak,
            "kept_frames": kept_l, "bpf": sum(10 * k for k in kept_l) / total, "loss_sum": float(loss.item()),
            "outputs": "codes int64, z_q, mask (no z_q_is, no latents)", "level": args.level,
            "data": "synthetic: per-item seeded N(0,1) latents / U(0,1) importance maps generated on the owning device",
            "collective": "none on the data path; one all_reduce of 9 numbers after the timed region (reduce_counts)",
        }
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
