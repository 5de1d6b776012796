"""Coordinate-range shards inside chromosomes (seeksv_b200/sharding.py: plan_range_shards, merge_range_texts,
sharded_getclip_ranges), CPU side: the plan is computed with the product's host-only ABI (.bai linear index, record peek,
virtual-offset distances) and the per-shard work is done by the CPU oracle on exactly the records the plan names, so what
is tested is the sharding theory itself - the halo the linear index promises, key ownership, the per-chromosome merge and
the cross-shard mate pairing - against the reference's whole-file outputs."""
import bisect
import os
import struct
import zlib

import pytest

from conftest import GOLDEN, read_text
from oracle import bamio, getclip_oracle
from seeksv_b200 import lib, mgpu, sharding

CASES = [("fuzz", "f11"), ("fuzz", "f12"), ("micro", "tumor"), ("example", "cancer"), ("example", "normal"), ("fuzz", "f106"), ("fuzz", "e3")]


def records_with_voffsets(path):
    raw = open(path, "rb").read()
    blocks, o, u, stream = [], 0, 0, bytearray()
    while o < len(raw):
        bsize = struct.unpack_from("<H", raw, o + 16)[0] + 1
        data = zlib.decompress(raw[o + 18:o + bsize - 8], -15)
        if data:
            blocks.append((o, u, len(data)))
        stream += data
        u += len(data)
        o += bsize
    h, recs, first = bamio.parse_bam_stream(bytes(stream))
    starts = [b[1] for b in blocks]
    voffs, p = [], first
    for r in recs:
        i = bisect.bisect_right(starts, p) - 1
        voffs.append(blocks[i][0] << 16 | (p - blocks[i][1]))
        p += r.size
    return h, recs, voffs


class OracleRangeWorker:
    """RangeShardWorker's protocol with the CPU oracle doing the work"""

    def __init__(self, h, recs, voffs, plan):
        self.h, self.plan = h, plan
        idx = {v: i for i, v in enumerate(voffs)}
        self.view, self.n_halo = [], 0
        if not plan.empty:
            end = len(recs) if plan.v_end is None else idx[plan.v_end]
            self.view = recs[(idx[plan.v_view] if plan.v_view else 0):end]
            self.n_halo = (idx[plan.v_own] if plan.v_own else 0) - (idx[plan.v_view] if plan.v_view else 0)
        if not plan.empty and plan.v_view:
            # the byte distances of the plan are the sizes of the records in between
            assert plan.halo_bytes == sum(r.size for r in self.view[:self.n_halo])
            n_ctx = idx[plan.v_halo] - idx[plan.v_view]
            assert plan.context_bytes == sum(r.size for r in self.view[:n_ctx])
            assert any(not (r.flag & 12) for r in self.view[:n_ctx]), "context without a mapped-branch record"

    def getclip(self):
        view = list(self.view)
        if self.plan.prev_tid != 0:   # the oracle starts last_tid at 0: a phantom mapped-branch record sets it (and is dropped itself)
            view = [bamio.make_rec("phantom", 0, self.plan.prev_tid, 0, 60, "10M", -1, -1, 0, "A" * 10, "I" * 10)] + view
        clip, fq, _, _ = getclip_oracle.getclip(self.h, view)
        lines, fql = clip.split("\n")[:-1], fq.split("\n")[:-1]
        keep, keepfq = [], []
        for i, line in enumerate(lines):
            f = line.split("\t", 3)
            if self.plan.key_lo <= (self.h.names.index(f[0]), int(f[1])) < self.plan.key_hi:
                keep.append(line)
                keepfq.extend(fql[4 * i:4 * i + 4])
        unmapped = b"".join(bamio.pack_record(r) for r in self.view[self.n_halo:] if r.flag & 12)
        return "".join(x + "\n" for x in keep), "".join(x + "\n" for x in keepfq), "", "", unmapped

    def pair_unmapped(self, records):
        _, urecs, _ = bamio.parse_bam_stream(bamio.header_bytes(self.h) + records)
        out = getclip_oracle.getclip(self.h, urecs)
        return out[2], out[3]


@pytest.mark.parametrize("d,s", CASES)
def test_range_shards_reproduce_the_whole_file(d, s, tmp_path):
    path = os.path.join(GOLDEN, d, s + ".sort.bam")
    h, recs, voffs = records_with_voffsets(path)
    golden = [read_text(os.path.join(GOLDEN, d, s + e)) for e in (".clip.txt", ".clip.fq.txt", ".unmapped_1.fq.txt", ".unmapped_2.fq.txt")]
    seen_multi = False
    for world in (1, 2, 3, 5, 8):
        plans = sharding.plan_range_shards(path, None, len(h.names), world)
        assert len(plans) == world
        live = [p for p in plans if not p.empty]
        seen_multi |= len(live) > 1
        # the own regions tile the file
        assert live[0].v_own == 0 and live[-1].v_end is None
        assert all(a.v_end == b.v_own and a.key_hi == b.key_lo for a, b in zip(live, live[1:]))
        workers = [OracleRangeWorker(h, recs, voffs, p) for p in plans]
        parts = [w.getclip() for w in workers]
        clip, fq = sharding.merge_range_texts([(p[0], p[1]) for p in parts])
        u1, u2 = workers[0].pair_unmapped(b"".join(p[4] for p in parts))
        assert (clip, fq, u1, u2) == tuple(golden), world
        # the multi-GPU commands' way (seeksv_b200/mgpu.py): every rank writes its blocks as gzip files, rank 0 orders the files
        prefix = str(tmp_path / ("w%d" % world))
        lists = [lib.write_range_blocks("%s.part%d" % (prefix, r), p[0].encode("latin-1"), p[1].encode("latin-1"), 2) for r, p in enumerate(parts)]
        mgpu.assemble_block_files(prefix, lists)
        assert lib.read_gz(prefix + ".clip.gz").decode("latin-1") == golden[0], world
        assert lib.read_gz(prefix + ".clip.fq.gz").decode("latin-1") == golden[1], world
        assert not [f for f in os.listdir(str(tmp_path)) if ".part" in f]
    assert seen_multi


def test_merge_orders_sides_per_chromosome(tmp_path):
    a = ("c1\t5\t5\tx\nc1\t9\t3\tx\n", "@a\nA\n+\nI\n@b\nC\n+\nI\n")
    b = ("c1\t20\t5\ty\nc1\t30\t3\ty\nc2\t4\t5\tz\n", "@c\nG\n+\nI\n@d\nT\n+\nI\n@e\nN\n+\nI\n")
    clip, fq = sharding.merge_range_texts([a, b])
    assert clip == "c1\t5\t5\tx\nc1\t20\t5\ty\nc1\t9\t3\tx\nc1\t30\t3\ty\nc2\t4\t5\tz\n"
    assert fq == "@a\nA\n+\nI\n@c\nG\n+\nI\n@b\nC\n+\nI\n@d\nT\n+\nI\n@e\nN\n+\nI\n"
    prefix = str(tmp_path / "m")
    lists = [lib.write_range_blocks("%s.part%d" % (prefix, r), p[0].encode(), p[1].encode(), 1) for r, p in enumerate((a, b))]
    assert lists == [[(b"c1", b"5"), (b"c1", b"3")], [(b"c1", b"5"), (b"c1", b"3"), (b"c2", b"5")]]
    mgpu.assemble_block_files(prefix, lists)
    assert lib.read_gz(prefix + ".clip.gz").decode() == clip and lib.read_gz(prefix + ".clip.fq.gz").decode() == fq
    # no blocks at all: empty, valid gzip files
    mgpu.assemble_block_files(prefix + "e", [[], []])
    assert lib.read_gz(prefix + "e.clip.gz") == b"" and lib.read_gz(prefix + "e.clip.fq.gz") == b""
