"""Batch x frame sharding of the RVQ hot path across the GPUs of one box.

Every op on the path is per-frame independent (all convs are 1x1, models/quantize.py:38-39), so a shard is a
set of (batch item, frame range) segments with no halo and no data-path collective.  One process per GPU; the
only cross-rank datum is the bits-per-frame numerator (per-stage kept-frame counts) and the loss sum, reduced on
the host with a tiny all_reduce over whatever backend the process group uses (NCCL on GPUs, gloo in CPU tests).
"""
from dataclasses import dataclass
from typing import List

TILE = 32  # kernel tile (frames); frame ranges are cut on tile boundaries so shards never split a tile


@dataclass(frozen=True)
class Segment:
    b: int   # batch item
    t0: int  # first frame (multiple of TILE)
    t1: int  # one past the last frame

    @property
    def frames(self):
        return self.t1 - self.t0


def plan_shards(B: int, T: int, world_size: int) -> List[List[Segment]]:
    """Split the B*ceil(T/TILE) tiles into `world_size` contiguous, near-equal runs (in (b, tile) order).

    Returns, per rank, the list of segments (a run may cover the tail of one item, whole items, and the head of
    another).  Ranks may get an empty list when there are fewer tiles than ranks."""
    if B < 0 or T < 0 or world_size < 1:
        raise ValueError("bad shard arguments")
    tpb = (T + TILE - 1) // TILE
    total = B * tpb
    out: List[List[Segment]] = []
    for r in range(world_size):
        lo = (total * r) // world_size
        hi = (total * (r + 1)) // world_size
        segs: List[Segment] = []
        i = lo
        while i < hi:
            b, tt = divmod(i, tpb)
            run = min(hi - i, tpb - tt)
            segs.append(Segment(b, tt * TILE, min(T, (tt + run) * TILE)))
            i += run
        out.append(segs)
    return out


def merge_whole_items(segs: List[Segment], T: int):
    """Group consecutive whole-item segments into (b0, b1) batch ranges so they run as one launch.
    Returns a list of ('items', b0, b1) and ('frames', b, t0, t1) work units."""
    units = []
    for s in segs:
        if s.t0 == 0 and s.t1 == T:
            if units and units[-1][0] == "items" and units[-1][2] == s.b:
                units[-1] = ("items", units[-1][1], s.b + 1)
            else:
                units.append(("items", s.b, s.b + 1))
        else:
            units.append(("frames", s.b, s.t0, s.t1))
    return units


def encode_shard(quantizer_weights, z, segs: List[Segment], n_run, imp_map=None, level=None, want_z_q_is=False):
    """Run the fused encode on this rank's segments of a (full-size, device-resident) latent z [B,D,T].
    Outputs are written in place into full-size tensors through strided views; returns an EncodeOutputs whose
    accumulators (loss sum, kept counts) cover exactly this rank's frames."""
    from . import ops

    B, D, T = z.shape
    out = ops.EncodeOutputs(B, D, T, n_run, z.device, z_q=True, z_q_is=want_z_q_is, latents=True, mask=True)
    frames = 0
    for u in merge_whole_items(segs, T):
        if u[0] == "items":
            bs, ts = slice(u[1], u[2]), slice(0, T)
        else:
            bs, ts = slice(u[1], u[1] + 1), slice(u[2], u[3])
        view = ops.EncodeOutputs.__new__(ops.EncodeOutputs)
        view.codes = out.codes[bs, :, ts]
        view.z_q = out.z_q[bs, :, ts]
        view.z_q_is = out.z_q_is[bs, :, :, ts] if out.z_q_is is not None else None
        view.latents = out.latents[bs, :, ts]
        view.mask = out.mask[bs, :, ts]
        view.loss_pf = None
        view.accum, view.n_run = out.accum, n_run
        lv = level
        if hasattr(level, "numel") and level.numel() == B:
            lv = level.reshape(-1)[bs]
        ops.rvq_encode_into(quantizer_weights, z[bs, :, ts], view, n_run,
                            None if imp_map is None else imp_map.reshape(B, T)[bs, ts], lv, zero_accum=False)
        frames += (bs.stop - bs.start) * (ts.stop - ts.start)
    out.frames = frames
    return out


def reduce_counts(kept, loss_sum, group=None):
    """Host-side reduction of the per-rank kept counts / loss sums (the only cross-rank exchange on the path)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(kept, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, group=group)
    return kept, loss_sum
