"""CPU: the multi-GPU plumbing (shard -> broadcast panel -> local compute -> gather) with world_size 2 over gloo.
The per-rank engine is replaced by the checker so that only the distribution logic is under test."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from gkl_b200 import multi, synth


def test_shard_bounds_are_contiguous_and_balanced():
    lens = np.random.default_rng(0).integers(35, 251, size=1000)
    for world in (1, 2, 3, 8):
        b = multi.shard_bounds(lens, world)
        assert b[0][0] == 0 and b[-1][1] == 1000 and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        loads = [lens[lo:hi].sum() for lo, hi in b]
        assert max(loads) - min(loads) <= 2 * 250
    assert multi.shard_bounds(np.array([100, 100]), 4)[-1][1] == 2  # more ranks than reads: some shards are empty


def _worker(rank, world, port, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = synth.random_batch(seed, 37, 9, unrelated=0.2) if rank == 0 else None
        eng = multi.ShardedPairHmm(lambda b: torch.from_numpy(oracle.port_pairhmm(b)[0]))
        out = eng.compute(batch, root=0)
        if rank == 0:
            q.put(out)
        # the bench path: every rank owns its reads, panel from rank 0, equal slabs gathered
        local = synth.config2(8, 4, seed=100 + rank)
        hap_off, hap = multi.broadcast_panel(local.hap_off if rank == 0 else None,
                                             torch.from_numpy(local.hap_bases) if rank == 0 else None, 0)
        res = torch.from_numpy(np.full(8 * (len(hap_off) - 1), float(rank)))
        g = multi.gather_slabs(res, [res.numel()] * world, 0)
        if rank == 0:
            q.put((hap_off, hap.numpy(), g.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_compute_equals_single_process(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, 77, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    hap_off, hap, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = synth.random_batch(77, 37, 9, unrelated=0.2)
    assert np.array_equal(out, oracle.port_pairhmm(full)[0])  # same bits, same r * H + h order
    ref = synth.config2(8, 4, seed=100)
    assert np.array_equal(hap_off, ref.hap_off) and np.array_equal(hap, ref.hap_bases)
    assert np.array_equal(gathered, np.repeat(np.arange(world, dtype=np.float64), 8 * 4))
