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


# ---- coordinate-range shards inside chromosomes ---------------------------------------------------------
# The .bai's linear index holds, for every 16 kb window of every reference, the virtual offset of the first record that
# overlaps the window. These offsets are record boundaries, so a BAM can be cut there without looking at the data, and they
# bound the reach of earlier records exactly: every record whose alignment ends beyond the start of window w lies at or
# after linear[w]. A shard that owns the records from cut c on therefore loads
#     [context][halo][own records ...)      context: from the previous index offset (gives the walk a predecessor record)
#                                           halo:    from linear[window(pos of the record at c) - 1] up to c
# and clusters only the breakpoint keys (tid, pos) >= (tid_c, pos_c + 1) and below the next shard's bound: every key is
# clustered on exactly one shard, with all of its reads, in file order. No data is exchanged between ranks; all ranks
# compute the same plan from the index. Outputs merge per chromosome: the '5' lines of all shards, then the '3' lines.
class RangePlan:
    def __init__(self, v_view, v_halo, v_own, v_end, key_lo, key_hi, prev_tid, halo_bytes, context_bytes):
        self.v_view, self.v_halo, self.v_own, self.v_end = v_view, v_halo, v_own, v_end
        self.key_lo, self.key_hi, self.prev_tid = key_lo, key_hi, prev_tid
        self.halo_bytes, self.context_bytes = halo_bytes, context_bytes

    @property
    def empty(self):
        return self.v_own is not None and self.v_own == self.v_end


KEY_MIN, KEY_MAX = (-(2 ** 31), -(2 ** 31)), (2 ** 31 - 1, 2 ** 31 - 1)


_PLAN_CACHE = {}


def plan_range_shards(bam_path: str, bai_path: Optional[str], n_ref: int, world: int, context_steps: int = 1) -> List[RangePlan]:
    """the same plan on every rank, from the .bai alone (memoised per file state: the commands of one process plan once)"""
    import os
    bai_path = bai_path or bam_path + ".bai"
    key = (os.path.abspath(bam_path), os.path.abspath(bai_path), os.path.getmtime(bam_path), os.path.getsize(bam_path),
           os.path.getmtime(bai_path), n_ref, world, context_steps)
    if key not in _PLAN_CACHE:
        if len(_PLAN_CACHE) > 64:
            _PLAN_CACHE.clear()
        _PLAN_CACHE[key] = _plan_range_shards(bam_path, bai_path, n_ref, world, context_steps)
    import copy
    return copy.deepcopy(_PLAN_CACHE[key])


def _plan_range_shards(bam_path: str, bai_path: str, n_ref: int, world: int, context_steps: int) -> List[RangePlan]:
    import bisect
    import os
    from . import lib
    lin = [lib.bai_linear_offsets(bai_path, t) for t in range(n_ref)]
    vset = sorted({v for L in lin for v in L if v})
    size = os.path.getsize(bam_path)
    cuts = []
    for r in range(1, world):
        if not vset:
            break
        target = size * r // world
        j = bisect.bisect_left(vset, target << 16)
        cand = [vset[k] for k in (j - 1, j) if 0 <= k < len(vset)]
        v = min(cand, key=lambda x: abs((x >> 16) - target))
        if v > vset[0]:
            cuts.append(v)
    cuts = sorted(set(cuts))
    bounds = [lib.peek_record(bam_path, c) for c in cuts]
    plans = []
    for k in range(len(cuts) + 1):
        v_own = cuts[k - 1] if k else 0
        v_end = cuts[k] if k < len(cuts) else None
        key_lo = (bounds[k - 1][0], bounds[k - 1][1] + 1) if k else KEY_MIN
        key_hi = (bounds[k][0], bounds[k][1] + 1) if k < len(cuts) else KEY_MAX
        if k == 0:
            plans.append(RangePlan(0, 0, 0, v_end, key_lo, key_hi, 0, 0, 0))
            continue
        tid, pos = bounds[k - 1]
        w = max((pos >> 14) - 1, 0)
        L = lin[tid]
        hv = 0
        for i in range(min(w, len(L) - 1), -1, -1):   # (a window nobody overlaps has no entry: the previous one bounds it)
            if L[i]:
                hv = L[i]
                break
        if not hv:
            hv = min(v for v in L if v)
        hv = min(hv, v_own)
        j = bisect.bisect_left(vset, hv)
        j_view = j - context_steps
        v_view = vset[j_view] if j_view >= 0 else 0
        prev_tid = lib.peek_record(bam_path, v_view)[0] if v_view else 0
        start = v_view or vset[0]
        plans.append(RangePlan(v_view, hv, v_own, v_end, key_lo, key_hi, prev_tid,
                               lib.voffset_distance(bam_path, start, v_own), lib.voffset_distance(bam_path, start, hv)))
    while len(plans) < world:   # more ranks than cut points: the rest get nothing
        plans.append(RangePlan(0, 0, 0, 0, KEY_MAX, KEY_MAX, 0, 0, 0))
        plans[-1].v_own = plans[-1].v_end = 1
    return plans


def merge_range_texts(parts: Sequence[Tuple[str, str]]) -> Tuple[str, str]:
    """parts: (clip text, clip.fq text) of the range shards in file order -> the whole-file texts: per chromosome (file order)
    the '5' lines of all shards, then the '3' lines (DisplaySClipReadsAndClipFq, clip_reads.h:300-345); 4 FASTQ lines per line"""
    order, groups = [], {}
    for clip, fq in parts:
        lines, fql = clip.split("\n")[:-1], fq.split("\n")[:-1]
        assert len(fql) == 4 * len(lines)
        for i, line in enumerate(lines):
            f = line.split("\t", 3)
            key = (f[0], f[2])
            if f[0] not in order:
                order.append(f[0])
            g = groups.setdefault(key, ([], []))
            g[0].append(line)
            g[1].extend(fql[4 * i:4 * i + 4])
    out, outfq = [], []
    for chrom in order:
        for side in ("5", "3"):
            g = groups.get((chrom, side))
            if g:
                out.extend(g[0])
                outfq.extend(g[1])
    return "".join(x + "\n" for x in out), "".join(x + "\n" for x in outfq)


class RangeShardWorker:
    """One rank's coordinate-range shard of an indexed BAM on the GPU (plan_range_shards)."""

    def __init__(self, ctx, bam_path: str, plan: RangePlan, **getclip_kw):
        from . import lib
        self.ctx, self.path, self.plan, self.kw = ctx, bam_path, plan, getclip_kw
        self.bam = lib.Bam.open_voffsets(ctx, bam_path, plan.v_view, plan.v_end) if not plan.empty else None

    def context_has_mapped_record(self) -> bool:
        """the walk needs a mapped-branch record in front of the halo (else plan again with a larger context_steps)"""
        if self.bam is None or not self.plan.v_view:
            return True
        from . import lib
        dptr, _, _ = self.bam.device_stream()
        ctx_view = lib.Bam.from_device(self.ctx, dptr, self.plan.context_bytes, 0, len(self.bam.ref_names))
        try:
            return ctx_view.last_mapped_tid() is not None
        finally:
            ctx_view.close()

    def getclip(self):
        if self.bam is None:
            return "", "", "", "", b""
        p = self.plan
        texts = self.bam.getclip(prev_tid=p.prev_tid, export_unmapped=True, key_range=(p.key_lo, p.key_hi), halo_bytes=p.halo_bytes,
                                 **self.kw)
        return tuple(t.decode("latin-1") for t in texts) + (self.bam.last_unmapped_records,)

    def pair_unmapped(self, records: bytes) -> Tuple[str, str]:
        return RefShardWorker.pair_unmapped(self, records)

    def own_view(self):
        """the shard's own records only (no context, no halo): what the additive getsv passes run on"""
        if self.bam is None:
            return None
        from . import lib
        dptr, nbytes, _ = self.bam.device_stream()
        v = lib.Bam.from_device(self.ctx, dptr + self.plan.halo_bytes, nbytes - self.plan.halo_bytes, 0, len(self.bam.ref_names),
                                keep=self.bam)
        v.set_refs(self.bam.ref_names, self.bam.ref_lens)
        return v

    def close(self):
        if self.bam is not None:
            self.bam.close()


def sharded_getclip_ranges(worker, dist=None) -> Optional[Tuple[str, str, str, str]]:
    """worker.getclip() -> (clip, clip.fq, "", "", unmapped-branch records) of its range shard; rank 0 merges"""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    parts = all_gather_objects(worker.getclip(), dist)
    if rank != 0:
        return None
    clip, fq = merge_range_texts([(p[0], p[1]) for p in parts])
    u1, u2 = worker.pair_unmapped(b"".join(p[4] for p in parts))
    return clip, fq, u1, u2


# ---- getsv / somatic on shards: additive device passes, combined with collectives on (device) tensors ------------------------------
# Every rank holds a `worker` over its own records (lib.Bam with set_own_offset for range shards; the CPU oracle in the gloo tests):
#   worker.insert_partial(min_mapq, take) -> (taken, sum, sum of squares, records above 46340)     additive
#   worker.insert_sq(min_mapq, take, mean) -> sum of (int32)((isize - mean)^2)                     additive (wrap-exact second pass)
#   worker.pairs_depth(min_mapq, mean, dev, times, junctions, windows) -> tensor int32 [n_j + n_pos] additive
# Collectives run on tensors of `device` (cuda with NCCL over NVLink, cpu with gloo); nothing is pickled.
def _all_gather_i64(values, dist, device):
    import torch
    t = torch.tensor(list(values), dtype=torch.int64, device=device)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [t.tolist()]
    out = torch.empty(dist.get_world_size() * t.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t) if device != "cpu" else dist.all_gather(list(out.view(dist.get_world_size(), -1).unbind(0)), t)
    return out.view(dist.get_world_size(), -1).tolist()


def _all_reduce_i64(values, dist, device):
    import torch
    t = torch.tensor(list(values), dtype=torch.int64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t.tolist()


def mean_dev(n: int, sx: int, sq: int) -> Tuple[int, int]:
    """cluster.cpp:72-80: integer mean, (int)sqrt of the double quotient; sq = sum of squared differences from that mean"""
    import math
    if n == 0:
        return 0, 0
    return sx // n, int(math.sqrt(float(sq) / float(n)))


def sharded_insert_stats(worker, dist, device, min_mapq: int, max_pairs: int):
    """CalculateInsertsizeDeviation over the shards of one BAM: (n, mean, dev) of the first max_pairs qualifying records of the WHOLE
    file. One all_gather of four integers per rank; a second small all_reduce only when the -n cut falls inside a shard; a third
    only when an insert size is large enough for the reference's int products to wrap."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    full = worker.insert_partial(min_mapq, -1)
    every = _all_gather_i64(full, dist, device)
    counts = [e[0] for e in every]
    takes = prefix_cutoffs(counts, max_pairs) if max_pairs > 0 else [0] * len(counts)
    if takes == counts:
        n, sx, sxx, big = (sum(e[k] for e in every) for k in range(4))
    else:
        take = takes[rank]
        mine = full if take == counts[rank] else ((0, 0, 0, 0) if take == 0 else worker.insert_partial(min_mapq, take))
        n, sx, sxx, big = _all_reduce_i64(mine, dist, device)
    if n == 0:
        return 0, 0, 0
    mean = sx // n
    mean = ((mean + 2 ** 31) % 2 ** 32) - 2 ** 31          # stored to an int (cluster.cpp:72)
    if big == 0:
        sq = sxx - 2 * mean * sx + n * mean * mean
    else:       # the reference multiplies two ints (cluster.cpp:77): redo the squares with its truncation
        own_take = -1 if takes == counts else takes[rank]
        mine = worker.insert_sq(min_mapq, own_take, mean) if own_take != 0 else 0
        sq = _all_reduce_i64([mine], dist, device)[0]
    return n, mean, mean_dev(n, mean * n, sq)[1]


PILEUP_CAP = 8000  # libbam's pileup buffer refuses reads above this many at one position (quirk Q12)


def cap_cut_by_shards(own, total, n_junctions: int):
    """The libbam pileup cap is global per position. Every shard emulates it on its own records, so the summed depth equals the
    whole-file depth unless a position that reaches the cap got reads from MORE THAN ONE shard (a > 8000x pile-up cut by a shard
    boundary). Returns a 0/1 tensor (on the tensors' device): 1 = this rank contributed part, not all, of such a position."""
    d_own, d_tot = own[n_junctions:], total[n_junctions:]
    return ((d_tot >= PILEUP_CAP) & (d_own > 0) & (d_own < d_tot)).any().to(total.dtype).reshape(1)


def sharded_pairs_depth(worker, dist, min_mapq, mean, dev, times, junctions, windows, cap_flag=None):
    """per-junction discordant-pair counts and per-position window depth, summed over the shards with ONE all_reduce of the
    tensor the workers filled (device memory under NCCL). A pile-up that reaches libbam's cap across a shard boundary cannot be
    reproduced by adding shards up: it is detected and refused loudly - RuntimeError here, or, when the caller passes cap_flag (a
    one-element tensor it checks itself, e.g. after a timed loop), OR-ed into that tensor without a synchronisation."""
    t = worker.pairs_depth(min_mapq, mean, dev, times, junctions, windows)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        own = t.clone()
        dist.all_reduce(t)
        flag = cap_cut_by_shards(own, t, len(junctions))
        if cap_flag is not None:
            cap_flag |= flag.to(cap_flag.dtype)
        else:
            dist.all_reduce(flag)
            if int(flag.item()):
                raise RuntimeError("a pile-up of %d or more reads is cut by a shard boundary: libbam's pileup cap (bam2depth.cpp:17-142 via "
                                   "bam_plp) is global per position and is not reproduced by range shards - shard --by chromosome or run "
                                   "the single-process command" % PILEUP_CAP)
    return t


class GpuShardWorker:
    """the getsv-side `worker` on a GPU: a lib.Bam over this rank's records (range shards: own_offset = halo bytes)"""

    def __init__(self, bam, device):
        self.bam, self.device = bam, device
        self._arrays = None

    def insert_partial(self, min_mapq, take):
        return self.bam.insert_partial(min_mapq, take)

    def insert_sq(self, min_mapq, take, mean):
        return self.bam.insert_sq(min_mapq, take, mean)

    def prepare(self, junctions, windows):
        """C arrays and the result tensor, built once for a junction / window list"""
        import ctypes as C
        import torch
        from . import lib
        nj, nw = len(junctions), len(windows)
        j_arr = (lib.Junction * max(nj, 1))(*[lib.Junction(ut, up, dt, dp, us.encode(), ds.encode(), b"") for ut, up, us, dt, dp, ds in junctions])
        w_arr = (lib.Window * max(nw, 1))(*[lib.Window(*w) for w in windows])
        n_pos = sum(w[2] - w[1] + 1 for w in windows)
        out = torch.zeros(max(nj + n_pos, 1), dtype=torch.int32, device=self.device)
        self._arrays = (j_arr, nj, w_arr, nw, n_pos, out)
        return out

    def pairs_depth(self, min_mapq, mean, dev, times, junctions, windows):
        from . import lib
        if self._arrays is None:
            self.prepare(junctions, windows)
        j_arr, nj, w_arr, nw, n_pos, out = self._arrays
        self.bam.pairs_depth_raw(lib.PairParams(min_mapq, mean, dev, times), j_arr, nj, w_arr, nw, out.data_ptr(), out.data_ptr() + 4 * nj)
        return out


def fnv1a64(name: bytes) -> int:
    """the name hash that groups exported unmapped-branch records (export_partitions; getclip.cu:export_group)"""
    h = 0xcbf29ce484222325
    for b in name:
        h = ((h ^ b) * 0x100000001b3) & 0xffffffffffffffff
    return h


def exchange_unmapped(clusters, dist, world: int, device):
    """All-to-all of the exported unmapped-branch records over device memory: rank r receives, from every shard in file order,
    the records whose name hashes to group r (mates share a name: every pair is decided on exactly one rank). Returns a uint8
    device tensor holding the received records in file order (plus readable padding)."""
    import torch
    dptr, n = clusters.export_device()
    parts = clusters.export_parts(world)
    send_sizes = [parts[r + 1] - parts[r] for r in range(world)]
    sizes = _all_gather_i64(send_sizes, dist, device)                     # sizes[src][dst]
    rank = dist.get_rank()
    recv_sizes = [sizes[src][rank] for src in range(world)]
    send = _as_tensor(dptr, max(n, 1), device)[:n]
    recv = torch.zeros(sum(recv_sizes) + 256, dtype=torch.uint8, device=device)
    dist.all_to_all_single(recv[:sum(recv_sizes)], send, recv_sizes, send_sizes)
    return recv, sum(recv_sizes)


class _DevMem:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _as_tensor(ptr: int, n: int, device):
    """a uint8 tensor over device memory owned by the library (no copy)"""
    import torch
    if not ptr:
        return torch.zeros(n, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevMem(ptr, n), device=device)


def merge_range_texts_fast(parts: Sequence[Tuple[bytes, bytes]]) -> Tuple[bytes, bytes]:
    """merge_range_texts on bytes with numpy: a shard's clip text is ordered by (chromosome, side, position), so the lines of one
    (chromosome, side) block are contiguous and the whole-file text is a concatenation of blocks - no line is parsed twice."""
    import numpy as np
    order, blocks = [], {}
    for clip, fq in parts:
        if not clip:
            continue
        buf = np.frombuffer(clip, dtype=np.uint8)
        nl = np.flatnonzero(buf == 10)
        starts = np.concatenate(([0], nl[:-1] + 1))
        tabs = np.flatnonzero(buf == 9)
        first_tab = tabs[np.searchsorted(tabs, starts)]                  # end of the chromosome name
        third = tabs[np.searchsorted(tabs, starts) + 1] + 1              # the side character follows the second tab
        side = buf[third]
        # a block changes where the name or the side changes: compare (name length, side) cheaply, confirm names at the boundaries
        name_len = first_tab - starts
        change = np.flatnonzero((side[1:] != side[:-1]) | (name_len[1:] != name_len[:-1])) + 1
        cand = np.concatenate(([0], change, [len(starts)]))
        # same length and side but another name: split further by comparing the name bytes of neighbouring lines inside a block
        bounds = [0]
        for a, b in zip(cand[:-1], cand[1:]):
            L = int(name_len[a])
            if b - a > 1 and L:
                names = buf[(starts[a:b, None] + np.arange(L)[None, :])]
                diff = np.flatnonzero((names[1:] != names[:-1]).any(axis=1)) + 1 + a
                bounds.extend(int(x) for x in diff)
            bounds.append(int(b))
        bounds = sorted(set(bounds))
        fbuf = np.frombuffer(fq, dtype=np.uint8)
        fnl = np.flatnonzero(fbuf == 10)
        for a, b in zip(bounds[:-1], bounds[1:]):
            if a == b:
                continue
            chrom = clip[int(starts[a]):int(first_tab[a])]
            sd = bytes([int(side[a])])
            c0, c1 = int(starts[a]), int(nl[b - 1]) + 1
            f0 = 0 if a == 0 else int(fnl[4 * a - 1]) + 1
            f1 = int(fnl[4 * b - 1]) + 1
            if chrom not in order:
                order.append(chrom)
            g = blocks.setdefault((chrom, sd), ([], []))
            g[0].append(clip[c0:c1])
            g[1].append(fq[f0:f1])
    out, outfq = [], []
    for chrom in order:
        for sd in (b"5", b"3"):
            g = blocks.get((chrom, sd))
            if g:
                out.extend(g[0])
                outfq.extend(g[1])
    return b"".join(out), b"".join(outfq)
