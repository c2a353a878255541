"""World-size-2 CPU (gloo) test of the multi-GPU plumbing: regions shard by rank with no
data-path collective; only the timing/count reductions bench.py performs go through
torch.distributed.  The per-rank 'scoring' here is the host-side grid sizing (no GPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

import mipgen_b200 as mg
from mipgen_b200 import panel, shard
from mipgen_b200.panel import Config


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = Config()
    genome = panel.lcg_genome(panel.genome_length_for(12, 300, cfg), 5)
    regions = panel.make_regions(genome, 12, 80, 300, cfg, 6)
    costs = [mg.config_grid_size(cfg, r) for r in regions]
    mine = shard.lpt_assign(costs, world)[rank]
    local = float(sum(costs[i] for i in mine))
    t = torch.tensor([local], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    owned = [None] * world
    dist.all_gather_object(owned, mine)
    if rank == 0:
        q.put((float(t.item()), float(ms.item()), owned, float(sum(costs))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_region_sharding_gloo():
    world, port = 2, _free_port()
    ctx = tmp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    total, ms, owned, want_total = q.get()
    assert total == want_total
    assert ms == 11.0  # max over ranks
    assert sorted(i for o in owned for i in o) == list(range(12))
