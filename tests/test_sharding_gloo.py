"""World-size-2 gloo runs (CPU) of the multi-GPU host logic in seeksv_b200/sharding.py: the per-shard work is done by
the CPU oracle, so what is tested is the sharding itself - chromosome ownership, the Q1 hand-over, the merge order,
and the additivity of getsv's per-junction counts / per-window depth / insert-size prefix."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT, read_text


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch
        from oracle import bamio, getclip_oracle, getsv_oracle as G
        from seeksv_b200 import sharding
        d, s = case
        path = os.path.join(GOLDEN, d, s + ".sort.bam")
        h, recs = bamio.read_bam(path)
        lo, hi = sharding.assign_chromosomes(h.lengths, world)[rank]
        # shard = the records of my chromosomes; tid -1 records (none in the fixtures) would go to the last rank
        mine = [r for r in recs if lo <= r.tid < hi or (r.tid < 0 and rank == world - 1)]

        class OracleWorker:
            def last_mapped_tid(self):
                t = None
                for r in mine:
                    if not (r.flag & 12):
                        t = r.tid
                return t

            def getclip(self, prev_tid):
                # the oracle starts last_tid at 0: emulate prev_tid by a phantom state (clip_reads.h:407)
                clip, fq, _, _ = _getclip_with_prev(getclip_oracle, h, mine, prev_tid)
                # mates are paired by name across the whole file: export the unmapped-branch records, rank 0 pairs them
                return clip, fq, "", "", b"".join(bamio.pack_record(r) for r in mine if r.flag & 12)

            def pair_unmapped(self, records):
                _, urecs, _ = bamio.parse_bam_stream(bamio.header_bytes(h) + records)
                out = getclip_oracle.getclip(h, urecs)
                assert out[0] == "" and out[1] == ""
                return out[2], out[3]
        merged = sharding.sharded_getclip(OracleWorker(), dist)
        # getsv side: counts and depth are owned by one rank each -> all_reduce(sum) reproduces the whole-file values
        clip_text = read_text(os.path.join(GOLDEN, d, s + ".clip.txt"))
        ch, ca = bamio.read_alignments(os.path.join(GOLDEN, d, s + ".clip.sam"))
        jm = G.JunctionMap()
        G.join_clip_alignments(G.parse_clip_text(clip_text), ch, ca, jm)
        G.merge_junction(jm, 50)
        whole = G.insert_size_stats(recs, 20, 5000000)
        qual = sum(1 for r in mine if r.mapq >= 20 and not G.is_hard_clip(r) and (r.flag & 1) and (r.flag & 2) and not (r.flag & 1024) and r.isize > 0)
        counts = sharding.all_gather_objects(qual, dist)
        for cap in (5000000, 100, 7):
            cut = sharding.prefix_cutoffs(counts, cap)[rank]
            part = [0, 0]
            n = 0
            for r in mine:
                if n == cut:
                    break
                if r.mapq >= 20 and not G.is_hard_clip(r) and (r.flag & 1) and (r.flag & 2) and not (r.flag & 1024) and r.isize > 0:
                    part[0] += 1
                    part[1] += r.isize
                    n += 1
            t = torch.tensor(part, dtype=torch.int64)
            dist.all_reduce(t)
            ref_n, ref_sum = 0, 0
            for r in recs:
                if ref_n == cap:
                    break
                if r.mapq >= 20 and not G.is_hard_clip(r) and (r.flag & 1) and (r.flag & 2) and not (r.flag & 1024) and r.isize > 0:
                    ref_n += 1
                    ref_sum += r.isize
            assert t.tolist() == [ref_n, ref_sum], (cap, t.tolist(), ref_n, ref_sum)
        mean, dev = whole
        local = torch.tensor([G.discordant_pairs(h, mine, k, 20, mean, dev, 4) for k in jm.keys], dtype=torch.int64)
        dist.all_reduce(local)
        want = [G.discordant_pairs(h, recs, k, 20, mean, dev, 4) for k in jm.keys]
        assert local.tolist() == want
        dep_mine = G.depth_arrays(h, mine, 20)
        dep_all = G.depth_arrays(h, recs, 20)
        for tid in range(len(h.names)):
            t = torch.from_numpy(dep_mine[tid].copy())
            dist.all_reduce(t)
            assert (t.numpy() == dep_all[tid]).all()
        if rank == 0:
            q.put(merged)
    finally:
        dist.destroy_process_group()


def _getclip_with_prev(getclip_oracle, h, recs, prev_tid):
    """oracle getclip over a shard whose stream position is `after a mapped-branch record of tid prev_tid`"""
    if prev_tid == 0:
        return getclip_oracle.getclip(h, recs)
    # prepend a harmless mapped-branch record of tid prev_tid... which itself would flush + be dropped (tid != 0):
    # exactly the state we need (last_tid == prev_tid, nothing emitted but an empty flush)
    from oracle import bamio
    phantom = bamio.make_rec("phantom", 0, prev_tid, 0, 60, "10M", -1, -1, 0, "A" * 10, "I" * 10)
    return getclip_oracle.getclip(h, [phantom] + list(recs))


@pytest.mark.parametrize("case", [("micro", "tumor"), ("example", "cancer"), ("fuzz", "f11"), ("fuzz", "e3")])   # fuzz: mates in different shards
def test_chromosome_sharding_world2(case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    d, s = case
    for got, name in zip(merged, (".clip.txt", ".clip.fq.txt", ".unmapped_1.fq.txt", ".unmapped_2.fq.txt")):
        assert got == read_text(os.path.join(GOLDEN, d, s + name)), name


def test_assign_and_prefix_helpers():
    from seeksv_b200 import sharding
    assert sharding.assign_chromosomes([100, 100, 100, 100], 2) == [(0, 2), (2, 4)]
    assert sharding.assign_chromosomes([24000, 16000, 3215], 2) == [(0, 1), (1, 3)]
    sh = sharding.assign_chromosomes([5] * 24, 8)
    assert sh[0][0] == 0 and sh[-1][1] == 24 and all(a[1] == b[0] for a, b in zip(sh, sh[1:])) and all(hi > lo for lo, hi in sh)
    assert sharding.prev_tids([0, None, 3, 4]) == [0, 0, 0, 3]
    assert sharding.prefix_cutoffs([10, 10, 10], 15) == [10, 5, 0]


def _range_worker(rank, world, port, case, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from seeksv_b200 import sharding
        from test_range_sharding import OracleRangeWorker, records_with_voffsets
        d, s = case
        path = os.path.join(GOLDEN, d, s + ".sort.bam")
        h, recs, voffs = records_with_voffsets(path)
        plan = sharding.plan_range_shards(path, None, len(h.names), world)[rank]     # every rank computes the same plan
        merged = sharding.sharded_getclip_ranges(OracleRangeWorker(h, recs, voffs, plan), dist)
        if rank == 0:
            q.put(merged)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [("fuzz", "f12"), ("example", "normal"), ("fuzz", "e3")])
def test_range_sharding_world2(case):
    """sharded_getclip_ranges over a real process group (gloo, 2 ranks): plan from the .bai on every rank, oracle workers, rank 0 merges"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_range_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    d, s = case
    for got, name in zip(merged, (".clip.txt", ".clip.fq.txt", ".unmapped_1.fq.txt", ".unmapped_2.fq.txt")):
        assert got == read_text(os.path.join(GOLDEN, d, s + name)), name
