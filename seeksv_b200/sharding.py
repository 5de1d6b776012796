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
    """worker.last_mapped_tid() -> Optional[int]; worker.getclip(prev_tid) -> 4 texts. Rank 0 returns the merged texts.

    Mates of the unmapped branch are paired by name across the whole file (clip_reads.h:172-219), so a worker may return
    its unmapped-branch RECORDS instead of the two FASTQ texts (a 5th element: packed BAM records in file order, texts 2
    and 3 empty); rank 0 then pairs the concatenation with worker.pair_unmapped(records) -> (unmapped_1, unmapped_2)."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    lasts = all_gather_objects(worker.last_mapped_tid(), dist)
    mine = worker.getclip(prev_tids(lasts)[rank])
    parts = all_gather_objects(mine, dist)
    if rank != 0:
        return None
    texts = ["".join(p[i] for p in parts) for i in range(4)]
    if any(len(p) > 4 for p in parts):
        texts[2], texts[3] = worker.pair_unmapped(b"".join(p[4] for p in parts if len(p) > 4))
    return tuple(texts)


def prefix_cutoffs(counts: Sequence[int], max_pairs: int) -> List[int]:
    """How many of each shard's qualifying pairs fall inside the first max_pairs of the whole file (cluster.cpp:48-70)."""
    out, left = [], max_pairs
    for c in counts:
        take = min(c, max(left, 0))
        out.append(take)
        left -= take
    return out


class RefShardWorker:
    """The `worker` of sharded_getclip on a GPU: one rank's chromosome shard of ONE indexed BAM, cut with the .bai at exact
    record boundaries (svb_bam_open_refs) - only the shard's BGZF blocks are read, uploaded and inflated."""

    def __init__(self, ctx, bam_path: str, tid_begin: int, tid_end: int, bai: Optional[str] = None, **getclip_kw):
        from . import lib
        self.bam = lib.Bam.open_refs(ctx, bam_path, tid_begin, tid_end, bai)
        self.kw = getclip_kw

    def last_mapped_tid(self) -> Optional[int]:
        return self.bam.last_mapped_tid()

    def getclip(self, prev_tid: int):
        texts = self.bam.getclip(prev_tid=prev_tid, export_unmapped=True, **self.kw)
        return tuple(t.decode("latin-1") for t in texts) + (self.bam.last_unmapped_records,)

    def pair_unmapped(self, records: bytes) -> Tuple[str, str]:
        """the merging rank: the shards' unmapped-branch records, concatenated in file order, as one small stream"""
        if not records:
            return "", ""
        from . import lib
        mini = lib.Bam.from_host(self.bam.ctx, records, 0, len(self.bam.ref_names))
        mini.set_refs(self.bam.ref_names, self.bam.ref_lens)
        try:
            texts = mini.getclip(**self.kw)
        finally:
            mini.close()
        return texts[2].decode("latin-1"), texts[3].decode("latin-1")

    def close(self):
        self.bam.close()


def open_ref_shard(ctx, bam_path: str, rank: int, world: int, bai: Optional[str] = None, **getclip_kw) -> RefShardWorker:
    """rank's share of the references of bam_path, balanced by reference length (assign_chromosomes)"""
    from . import lib
    probe = lib.Bam.open_refs(ctx, bam_path, 0, 0, bai)   # header only: no records are loaded for an empty range
    lens = probe.ref_lens
    probe.close()
    lo, hi = assign_chromosomes(lens, world)[rank]
    return RefShardWorker(ctx, bam_path, lo, hi, bai, **getclip_kw)
