"""Multi-GPU plumbing: the genome is partitioned by chromosome across the ranks of one node.

The reference is a single-threaded process; sharding is ours. What makes chromosome shards exact:
  * getclip flushes its cluster maps at every chromosome switch (clip_reads.h:423-438), so clusters never span
    chromosomes and the whole-file clip.gz / clip.fq.gz are the per-shard outputs concatenated in tid order;
  * quirk Q1 (the first mapped-branch record after a chromosome switch is dropped) depends only on the tid of the
    last mapped-branch record BEFORE the shard - one integer per rank, exchanged with one all_gather;
  * getsv's discordant-pair query of a junction reads records of its up-chromosome only (getsv.cpp:1039-1067) and
    depth windows live on one chromosome, so per-junction counts and per-window depth are owned by exactly one rank
    and the merge is a sum (all_reduce) / gather of disjoint pieces; insert-size statistics use the first -n
    qualifying records in file order, i.e. a prefix over the shard counts.

Everything here is host logic over torch.distributed (NCCL on GPUs, gloo in the CPU tests); the per-shard work is
done by a `worker` object - the CUDA library on a GPU box, the CPU oracle in tests/test_sharding_gloo.py.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def assign_chromosomes(ref_lens: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous tid ranges [lo, hi) per rank, balanced by reference length (greedy prefix split)."""
    n = len(ref_lens)
    total = float(sum(ref_lens)) or 1.0
    out, lo, acc = [], 0, 0.0
    for r in range(world):
        hi = lo
        target = total * (r + 1) / world
        while hi < n and (acc + ref_lens[hi] / 2.0 <= target or hi == lo) and (n - hi) > (world - 1 - r):
            acc += ref_lens[hi]
            hi += 1
        if r == world - 1:
            hi = n
        out.append((lo, hi))
        lo = hi
    return out


def prev_tids(last_mapped_tid_per_rank: Sequence[Optional[int]]) -> List[int]:
    """prev_tid parameter of svb_getclip for every rank: tid of the last mapped-branch record of the closest earlier
    shard that has one; 0 for the first (clip_reads.h:407 starts last_tid at 0)."""
    out, cur = [], 0
    for t in last_mapped_tid_per_rank:
        out.append(cur)
        if t is not None:
            cur = t
    return out


def all_gather_objects(obj, dist=None):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def sharded_getclip(worker, dist=None) -> Optional[Tuple[str, str, str, str]]:
    """worker.last_mapped_tid() -> Optional[int]; worker.getclip(prev_tid) -> 4 texts. Rank 0 returns the merged texts."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    lasts = all_gather_objects(worker.last_mapped_tid(), dist)
    mine = worker.getclip(prev_tids(lasts)[rank])
    parts = all_gather_objects(mine, dist)
    if rank != 0:
        return None
    return tuple("".join(p[i] for p in parts) for i in range(4))


def prefix_cutoffs(counts: Sequence[int], max_pairs: int) -> List[int]:
    """How many of each shard's qualifying pairs fall inside the first max_pairs of the whole file (cluster.cpp:48-70)."""
    out, left = [], max_pairs
    for c in counts:
        take = min(c, max(left, 0))
        out.append(take)
        left -= take
    return out
