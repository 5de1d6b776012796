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


def _stats_worker(rank, world, port, q):
    """sharded_insert_stats / sharded_pairs_depth over gloo with synthetic per-shard insert sizes (no BAM involved): the whole-file
    rule - first max_pairs qualifying records in file order, integer mean, int products that may wrap - must come out"""
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import math
        import random
        import torch
        from seeksv_b200 import sharding
        results = []
        for case, (sizes, big, cap) in enumerate([((700, 900, 400), False, 5000000), ((700, 900, 400), False, 1000), ((5, 0, 9), False, 7),
                                                  ((300, 300, 300), True, 5000000), ((300, 300, 300), True, 450), ((0, 0, 0), False, 10)]):
            rng = random.Random(100 + case)
            shards = [[rng.randrange(200, 900) for _ in range(n)] for n in sizes]
            if big:
                shards[1][5] = 70000          # (isize - mean)^2 wraps an int
                shards[2][7] = 71000
            shards = (shards + [[]] * world)[:world] if world <= 3 else shards + [[]] * (world - 3)
            mine = shards[rank] if rank < len(shards) else []

            class W:
                def insert_partial(self, mq, take):
                    xs = mine if take < 0 else mine[:take]
                    return (len(xs), sum(xs), sum(x * x for x in xs), sum(1 for x in xs if x > 46340))

                def insert_sq(self, mq, take, mean):
                    xs = mine if take < 0 else mine[:take]
                    tot = 0
                    for x in xs:
                        p = ((x - mean) * (x - mean)) & 0xffffffff
                        tot += p - (1 << 32) if p & 0x80000000 else p
                    return tot

                def pairs_depth(self, mq, mean, dev, times, juncs, wins):
                    return torch.tensor([rank + 1, 10 * (rank + 1), mean], dtype=torch.int32)
            n, mean, dev = sharding.sharded_insert_stats(W(), dist, "cpu", 20, cap)
            flat = [x for s_ in shards for x in s_][:cap]
            if not flat:
                want = (0, 0, 0)
            else:
                m = sum(flat) // len(flat)
                sq = 0
                for x in flat:
                    p = ((x - m) * (x - m)) & 0xffffffff
                    sq += p - (1 << 32) if p & 0x80000000 else p
                want = (len(flat), m, int(math.sqrt(float(sq) / float(len(flat)))))
            results.append(((n, mean, dev), want))
            t = sharding.sharded_pairs_depth(W(), dist, 20, mean, dev, 4, [], [])
            results.append((t.tolist()[:2], [sum(r + 1 for r in range(world)), sum(10 * (r + 1) for r in range(world))]))

        # libbam's pileup cap is global per position: a >= 8000x position fed by TWO shards must be refused loudly, on every rank;
        # the same depth from one shard alone is fine (that shard's own cap emulation is exact)
        class C:
            def __init__(self, vals):
                self.vals = vals

            def pairs_depth(self, *a):
                return torch.tensor(self.vals, dtype=torch.int32)
        juncs, wins = [(0, 1, "+", 0, 2, "-")], [(0, 1, 3)]        # one junction count in front of three depth positions
        cut = {0: [3, 5000, 10, 0], 1: [2, 4000, 0, 7]}.get(rank, [0, 0, 0, 0])
        try:
            sharding.sharded_pairs_depth(C(cut), dist, 20, 0, 0, 4, juncs, wins)
            raised = False
        except RuntimeError as e:
            raised = "pile-up" in str(e)
        results.append(([raised], [True]))
        flag = torch.zeros(1, dtype=torch.int32)
        sharding.sharded_pairs_depth(C(cut), dist, 20, 0, 0, 4, juncs, wins, cap_flag=flag)        # the caller checks later (bench.py)
        results.append(([int(flag.item())], [1 if rank < 2 else 0]))
        whole = {0: [3, 9000, 10, 0], 1: [2, 0, 0, 7]}.get(rank, [0, 0, 0, 0])
        t = sharding.sharded_pairs_depth(C(whole), dist, 20, 0, 0, 4, juncs, wins)
        results.append((t.tolist(), [5, 9000, 10, 7]))
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_statistics_and_sums_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stats_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, results in got:
        for have, want in results:
            assert tuple(have) == tuple(want), (rank, have, want)


def test_text_merge_fast_equals_reference_merge():
    from seeksv_b200 import sharding
    for d, s in (("micro", "tumor"), ("example", "cancer"), ("fuzz", "f11"), ("fuzz", "f12")):
        clip = read_text(os.path.join(GOLDEN, d, s + ".clip.txt"))
        fq = read_text(os.path.join(GOLDEN, d, s + ".clip.fq.txt"))
        lines, fl = clip.split("\n")[:-1], fq.split("\n")[:-1]
        for cuts in ((), (len(lines) // 3,), (len(lines) // 4, len(lines) // 2, len(lines) // 2 + 1)):
            b = [0] + list(cuts) + [len(lines)]
            parts = [("".join(x + "\n" for x in lines[i:j]), "".join(x + "\n" for x in fl[4 * i:4 * j])) for i, j in zip(b[:-1], b[1:])]
            want = sharding.merge_range_texts(parts)
            got = sharding.merge_range_texts_fast([(p[0].encode("latin-1"), p[1].encode("latin-1")) for p in parts])
            assert got[0] == want[0].encode("latin-1") and got[1] == want[1].encode("latin-1"), (d, s, cuts)
