"""CPU: batch x frame shard planning and the world_size-2 host-side reduction (gloo).  The per-rank compute in
the gloo test is the oracle (test infrastructure) -- what is checked is the host logic of the N>1 path:
disjoint cover, tile alignment, concatenation == full result, and the kept-count / loss all_reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vrvq_b200 import sharding
from tests.golden import gen_inputs as gi


@pytest.mark.parametrize("B,T,world", [(16, 862, 1), (16, 862, 2), (16, 862, 8), (256, 5168, 8), (1, 87, 8), (3, 100, 4), (5, 31, 2), (0, 10, 2), (2, 0, 2)])
def test_plan_is_a_disjoint_tile_aligned_cover(B, T, world):
    plan = sharding.plan_shards(B, T, world)
    assert len(plan) == world
    seen = np.zeros((B, T), np.int32)
    for segs in plan:
        for s in segs:
            assert s.t0 % sharding.TILE == 0 and (s.t1 % sharding.TILE == 0 or s.t1 == T) and s.t0 < s.t1
            seen[s.b, s.t0:s.t1] += 1
    assert (seen == 1).all()
    tiles = [sum((s.frames + sharding.TILE - 1) // sharding.TILE for s in segs) for segs in plan]
    assert max(tiles) - min(tiles) <= 1, "shards must be balanced to within one tile"


def test_merge_whole_items():
    plan = sharding.plan_shards(8, 64, 2)
    units = sharding.merge_whole_items(plan[0], 64)
    assert units == [("items", 0, 4)]
    plan = sharding.plan_shards(1, 320, 4)
    assert sharding.merge_whole_items(plan[1], 320) == [("frames", 0, 64, 160)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle

    torch.set_num_threads(1)
    c_oracle.set_num_threads(1)
    sd = gi.torch_state_dict(gi.make_state_dict(61, 4, 256))
    w = c_oracle.OracleWeights.from_state_dict(sd)
    z = gi.make_latents(62, B, 256, T, 1.0)
    imp = gi.make_imp_map(63, B, T)
    segs = sharding.plan_shards(B, T, world)[rank]
    codes = np.full((B, 4, T), -1, np.int64)
    kept = torch.zeros(4, dtype=torch.int64)
    loss = torch.zeros(1, dtype=torch.float64)
    for s in segs:
        o = c_oracle.encode(w, z[s.b:s.b + 1, :, s.t0:s.t1], None, imp[s.b:s.b + 1, :, s.t0:s.t1], 0.8, want_z_q_is=False)
        codes[s.b, :, s.t0:s.t1] = o["codes"][0]
        kept += torch.from_numpy(o["kept"])
        loss += o["loss_masked_sum"]
    sharding.reduce_counts(kept, loss)  # the only cross-rank exchange on the path
    gathered = [torch.empty_like(torch.from_numpy(codes)) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(codes))
    if rank == 0:
        merged = np.maximum.reduce([g.numpy() for g in gathered])
        full = c_oracle.encode(w, z, None, imp, 0.8, want_z_q_is=False)
        q.put((np.array_equal(merged, full["codes"]), kept.tolist() == full["kept"].tolist(),
               abs(loss.item() - full["loss_masked_sum"]) <= 1e-9 * abs(full["loss_masked_sum"])))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_shards_reduce_to_the_full_result():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 3, 100, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == (True, True, True), res
