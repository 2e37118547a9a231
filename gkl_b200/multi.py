"""Multi-GPU PairHMM: one process per GPU, reads sharded, haplotype panel broadcast, likelihood slabs gathered.

Every (read, haplotype) pair is independent (the reference exploits exactly that with its OpenMP loop,
pairhmm/IntelPairHmm.cc:151-169) and the output index is read-major (pairhmm/JavaData.h:94-105), so giving rank g
a contiguous block of reads gives it a contiguous slab of the output: no data-path exchange is needed beyond
distributing the panel and collecting the slabs.  The collectives are torch.distributed's (NCCL over NVLink on
GPUs; gloo in the CPU tests of this plumbing).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

from .batch import PairHmmBatch


def shard_bounds(read_lens: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous read ranges, one per rank, balanced by total read length (cells per read are
    proportional to its length: every read meets every haplotype)."""
    n = len(read_lens)
    cum = np.concatenate([[0], np.cumsum(read_lens, dtype=np.int64)])
    total = int(cum[-1])
    cuts = [0]
    for g in range(1, world):
        target = total * g / world
        i = int(np.searchsorted(cum, target, side="left"))
        if i > 0 and abs(cum[i - 1] - target) <= abs(cum[min(i, n)] - target):
            i -= 1
        cuts.append(min(max(i, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


def _dev(device) -> torch.device:
    return torch.device(device) if device is not None else torch.device("cpu")


def broadcast_panel(hap_off: Optional[np.ndarray], hap_bases: Optional[torch.Tensor], src: int, device=None,
                    group=None) -> tuple[np.ndarray, torch.Tensor]:
    """The haplotype panel (offsets + bases) travels from rank `src` to every rank."""
    dev = _dev(device)
    rank = dist.get_rank(group)
    meta = torch.zeros(2, dtype=torch.int64, device=dev)
    if rank == src:
        meta[0], meta[1] = len(hap_off), int(hap_off[-1])
    dist.broadcast(meta, src=src, group=group)
    n_off, n_bytes = int(meta[0]), int(meta[1])
    off_t = torch.from_numpy(hap_off.copy()).to(dev) if rank == src else torch.empty(n_off, dtype=torch.int64, device=dev)
    dist.broadcast(off_t, src=src, group=group)
    if rank == src:
        bases_t = hap_bases.to(dev)
    else:
        bases_t = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    dist.broadcast(bases_t, src=src, group=group)
    return off_t.cpu().numpy(), bases_t


def gather_slabs(local: torch.Tensor, counts: list[int], dst: int, group=None) -> Optional[torch.Tensor]:
    """Likelihood slabs (double[reads_g * H]) of all ranks, concatenated in rank order on `dst`.
    Equal slabs use one gather; unequal ones are padded to the longest."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    longest = max(counts)
    if local.numel() < longest:
        pad = torch.zeros(longest, dtype=local.dtype, device=local.device)
        pad[:local.numel()] = local
        local = pad
    bufs = [torch.empty(longest, dtype=local.dtype, device=local.device) for _ in range(world)] if rank == dst else None
    dist.gather(local, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])


class ShardedPairHmm:
    """computeLikelihoods over all ranks of a process group.

    compute_local(batch) -> torch.Tensor (double[reads * haps], on `device`) is the per-rank engine: in the
    product it wraps gkl_b200.native.Engine; the CPU tests of the plumbing inject a checker instead."""

    def __init__(self, compute_local: Callable[[PairHmmBatch], torch.Tensor], device=None, group=None):
        self.compute_local, self.device, self.group = compute_local, device, group

    def compute(self, batch: Optional[PairHmmBatch], root: int = 0) -> Optional[np.ndarray]:
        """`batch` is only read on `root`; the result (double[R * H]) is returned on `root`."""
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        dev = _dev(self.device)
        # 1. plan on root, tell everyone the shard sizes
        if rank == root:
            bounds = shard_bounds(batch.read_lens, world)
            plan = [(lo, hi, int(batch.read_off[lo]), int(batch.read_off[hi])) for lo, hi in bounds]
            obj = [plan, batch.n_haps]
        else:
            obj = [None, None]
        dist.broadcast_object_list(obj, src=root, group=self.group)
        plan, n_haps = obj
        lo, hi, b0, b1 = plan[rank]
        # 2. panel to everyone
        hap_off, hap_t = broadcast_panel(batch.hap_off if rank == root else None,
                                         torch.from_numpy(batch.hap_bases) if rank == root else None, root, dev,
                                         self.group)
        # 3. read shards to their owners (offsets + five arenas), padded to the longest shard
        max_bytes = max(p[3] - p[2] for p in plan)
        max_reads = max(p[1] - p[0] for p in plan)
        mine_off = torch.empty(max_reads + 1, dtype=torch.int64, device=dev)
        if rank == root:
            lists = []
            for (l, h, a, _) in plan:
                t = torch.zeros(max_reads + 1, dtype=torch.int64)
                t[:h - l + 1] = torch.from_numpy(batch.read_off[l:h + 1] - a)
                lists.append(t.to(dev))
        dist.scatter(mine_off, lists if rank == root else None, src=root, group=self.group)
        arenas = []
        for name in ("read_bases", "read_quals", "ins_gop", "del_gop", "gcp"):
            mine = torch.empty(max(1, max_bytes), dtype=torch.uint8, device=dev)
            if rank == root:
                src_arr = getattr(batch, name)
                lists = []
                for (_, _, a, b) in plan:
                    t = torch.zeros(max(1, max_bytes), dtype=torch.uint8)
                    t[:b - a] = torch.from_numpy(src_arr[a:b])
                    lists.append(t.to(dev))
            dist.scatter(mine, lists if rank == root else None, src=root, group=self.group)
            arenas.append(mine.cpu().numpy()[:b1 - b0].copy())
        off = mine_off.cpu().numpy()[:hi - lo + 1].copy()
        # 4. local sweep, 5. slabs back to root
        if hi > lo:
            local = PairHmmBatch(off, *arenas, hap_off, hap_t.cpu().numpy())
            res = self.compute_local(local).to(dev)
        else:
            res = torch.empty(0, dtype=torch.float64, device=dev)
        out = gather_slabs(res, [(p[1] - p[0]) * n_haps for p in plan], root, self.group)
        return out.cpu().numpy() if rank == root else None
