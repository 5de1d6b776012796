"""The CPU oracle (oracle/*.py) against the reference's own outputs (tests/golden/, produced by
oracle/_ref/seeksv == reference v1.2.3, see tests/golden/make_golden.py). CPU only."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT, read_text
from oracle import bamio, getclip_oracle, getsv_oracle

GETSV_CASES = [("example", "cancer"), ("example", "normal"), ("micro", "tumor"), ("micro", "normal"),
               ("fuzz", "f11"), ("fuzz", "f12"), ("fuzz", "f106"), ("fuzz", "e3"), ("long", "lq")]   # fuzz: tests/fuzzgen.py through the reference binary
# (make_golden.py); f106 was found by tools/fuzz_campaign.py: a junction position in front of the first flank range of the
# smallest-named chromosome keeps point depth 0 (the `continue` at bam2depth.cpp:102 skips the store at :123-124);
# long/lq: clipped sequences of 254-600 bases as clip.sam read names (libbam's 8-bit l_qname shifts the join, make_golden.py)
CASES = GETSV_CASES + [("kat", "quirks"), ("kat", "start_tid1")]


def _bam(d, s):
    p = os.path.join(GOLDEN, d, s + ".sort.bam")
    return p if os.path.exists(p) else os.path.join(GOLDEN, d, s + ".bam")


@pytest.mark.parametrize("d,s", CASES)
def test_getclip_matches_reference(d, s):
    h, recs = bamio.read_bam(_bam(d, s))
    clip, fq, u1, u2 = getclip_oracle.getclip(h, recs)
    assert clip == read_text(os.path.join(GOLDEN, d, s + ".clip.txt"))
    assert fq == read_text(os.path.join(GOLDEN, d, s + ".clip.fq.txt"))
    assert u1 == read_text(os.path.join(GOLDEN, d, s + ".unmapped_1.fq.txt"))
    assert u2 == read_text(os.path.join(GOLDEN, d, s + ".unmapped_2.fq.txt"))


@pytest.mark.parametrize("d,s", GETSV_CASES)
def test_getsv_matches_reference(d, s):
    h, recs = bamio.read_bam(_bam(d, s))
    ch, ca = bamio.read_alignments(os.path.join(GOLDEN, d, s + ".clip.sam"))
    sv, out = getsv_oracle.getsv(h, recs, read_text(os.path.join(GOLDEN, d, s + ".clip.txt")), ch, ca)
    assert sv == read_text(os.path.join(GOLDEN, d, s + ".sv"))
    assert out == read_text(os.path.join(GOLDEN, d, s + ".getsv.stdout"))


@pytest.mark.parametrize("d,normal,tumour", [("example", "normal", "cancer"), ("micro", "normal", "tumor"),
                                             ("fuzz", "f11", "f11"), ("fuzz", "f12", "f12"), ("fuzz", "f106", "f106"),
                                             ("fuzz", "e3", "e3"), ("long", "lq", "lq")])
def test_somatic_matches_reference(d, normal, tumour):
    h, recs = bamio.read_bam(_bam(d, normal))
    got = getsv_oracle.somatic(h, recs, read_text(os.path.join(GOLDEN, d, normal + ".clip.txt")),
                               read_text(os.path.join(GOLDEN, d, tumour + ".sv")))
    assert got == read_text(os.path.join(GOLDEN, d, tumour + ".somatic.temp.sv"))


def test_insert_size_example():
    # reference stderr for the example: "Mean insert size : 500 / Mean deviation: 25" 
    for s in ("cancer", "normal"):
        h, recs = bamio.read_bam(_bam("example", s))
        assert getsv_oracle.insert_size_stats(recs, 20, 5000000) == (500, 25)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "bamtool")), reason="no oracle/_ref")
def test_aux_walk_against_libbam(tmp_path):
    """bamio.aux_get_int == bam_aux2i(bam_aux_get()) of the linked libbam on random aux blocks, including the float / double
    fields that library cannot step over (an XC behind one is normally lost - the reference then treats the read as XC=0)."""
    import random
    import struct
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fuzzgen
    rng = random.Random(3)
    h = bamio.Header(["c"], [1000], "")
    recs = []
    for i in range(3000):
        while True:
            parts = [fuzzgen._aux(rng, None)]
            if rng.random() < 0.7:
                parts.insert(rng.randrange(2), b"FLf" + struct.pack("<f", rng.random() * 1000))
            if rng.random() < 0.3:
                parts.insert(rng.randrange(len(parts) + 1), b"DBd" + struct.pack("<d", rng.random()))
            kind = rng.choice("cCsSiIAZ")
            val = {"c": struct.pack("<b", -7), "C": b"\xc8", "s": struct.pack("<h", -300), "S": struct.pack("<H", 60000),
                   "i": struct.pack("<i", rng.randrange(1, 300)), "I": struct.pack("<I", 4000000000), "A": b"7", "Z": b"12\0"}[kind]
            parts.insert(rng.randrange(len(parts) + 1), b"XC" + kind.encode() + val)
            aux = b"".join(parts)
            if not bamio.aux_walk(aux, b"XC")[1]:   # a walk that leaves the record reads stale buffer bytes in the library
                break
        recs.append(bamio.make_rec("r%d" % i, 0, 0, 10, 30, "10M", -1, -1, 0, "ACGTACGTAC", "IIIIIIIIII", aux))
    path = str(tmp_path / "aux.bam")
    bamio.write_bam(path, h, recs)
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "bamtool"), "auxi", path, "XC"], capture_output=True).stdout.split(b"\n")
    assert len(out) == len(recs) + 1
    lost = 0
    for line, r in zip(out, recs):
        want = int(line.rsplit(b"\t", 2)[2])
        assert bamio.aux_get_int(r.aux, b"XC") == want, (line, r.aux)
        lost += want == 0
    assert 0 < lost < len(recs)   # both outcomes are exercised


def test_calend_counts_m_d_n_only():
    # probed on the linked libbam: `oracle/_ref/bamtool calend 100 <cigar>`
    for cig, want in (("10M", 110), ("10M5D", 115), ("10M5N", 115), ("10M5I", 110), ("5S10M", 110),
                      ("10M5=", 110), ("10M5X", 110), ("10M5P", 110)):
        r = bamio.make_rec("q", 0, 0, 100, 60, cig, -1, -1, 0, "", "")
        assert getsv_oracle.calend(r) == want
        tool = os.path.join(ROOT, "oracle", "_ref", "bamtool")
        if os.path.exists(tool):
            assert int(subprocess.run([tool, "calend", "100", cig], capture_output=True, text=True).stdout) == want


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "bamtool")), reason="no oracle/_ref")
def test_pileup_depth_and_cap_against_libbam(tmp_path):
    """depth_arrays (incl. the 8000-read cap and the =/X quirk) == the linked libbam's own pileup."""
    import random
    rng = random.Random(3)
    h = bamio.Header(["c1", "c2"], [5000, 5000], "@SQ\tSN:c1\tLN:5000\n@SQ\tSN:c2\tLN:5000\n")
    recs = []
    for tid in (0, 1):
        pos = 10
        for block in range(40):
            pos += rng.choice((0, 0, 1, 3, 7, 60))
            n = rng.choice((1, 5, 200, 3000, 9000)) if tid == 0 else rng.choice((1, 2, 50))
            for i in range(n):
                cig = rng.choice(("50M", "20M3D27M", "10S40M", "25M2I23M", "10=5X30M", "30M10N10M", "45M5H"))
                flag = rng.choice((0, 0, 0, 16, 1024, 256, 4, 512))
                recs.append(bamio.make_rec("r", flag, tid, pos, rng.choice((0, 30, 60)), cig, -1, -1, 0, "A" * 50,
                                           "I" * 50))
    path = str(tmp_path / "cap.bam")
    bamio.write_bam(path, h, recs)
    tool = os.path.join(ROOT, "oracle", "_ref", "bamtool")
    for mq in (0, 20):
        want = {}
        for line in subprocess.run([tool, "depth", path, str(mq)], capture_output=True, text=True).stdout.splitlines():
            t, p, n, m = map(int, line.split("\t"))
            want[(t, p)] = n - m
        got = getsv_oracle.depth_arrays(h, recs, mq)
        for tid in (0, 1):
            for p in range(1, 5001):
                assert int(got[tid][p]) == want.get((tid, p), 0), (mq, tid, p)


@pytest.mark.parametrize("d,s", [("micro", "tumor"), ("fuzz", "f11"), ("example", "cancer")])
def test_getsv_with_seed_file_matches_reference(d, s):
    """getsv -B <an earlier output>: ReadBreakpoint (getsv.cpp:1291-1323) seeds the junction map before the join"""
    h, recs = bamio.read_bam(_bam(d, s))
    ch, ca = bamio.read_alignments(os.path.join(GOLDEN, d, s + ".clip.sam"))
    clip = read_text(os.path.join(GOLDEN, d, s + ".clip.txt"))
    seed = read_text(os.path.join(GOLDEN, d, s + ".sv"))
    sv, out = getsv_oracle.getsv(h, recs, clip, ch, ca, seed_text=seed)
    assert sv == read_text(os.path.join(GOLDEN, d, s + ".B.sv"))
    assert out == read_text(os.path.join(GOLDEN, d, s + ".B.stdout"))
    sv, out = getsv_oracle.getsv(h, recs, clip, ch, ca, seed_text=seed, pairs_used=0, output_depth=False)
    assert sv == read_text(os.path.join(GOLDEN, d, s + ".B.n0D.sv"))
    assert out == read_text(os.path.join(GOLDEN, d, s + ".B.n0D.stdout"))


@pytest.mark.parametrize("d,s", [("micro", "tumor"), ("fuzz", "f11")])
def test_getsv_with_connected_reads_matches_reference(d, s):
    """getsv -F <connected read-through reads>: FindJunction (process_bwasw.cpp:5-227), alone and together with -B / -w"""
    h, recs = bamio.read_bam(_bam(d, s))
    ch, ca = bamio.read_alignments(os.path.join(GOLDEN, d, s + ".clip.sam"))
    clip = read_text(os.path.join(GOLDEN, d, s + ".clip.txt"))
    connect = bamio.read_alignments(os.path.join(GOLDEN, d, s + ".connect.sam"))
    sv, out = getsv_oracle.getsv(h, recs, clip, ch, ca, connect=connect)
    assert out == read_text(os.path.join(GOLDEN, d, s + ".F.stdout"))
    assert sv == read_text(os.path.join(GOLDEN, d, s + ".F.sv"))
    sv, out = getsv_oracle.getsv(h, recs, clip, ch, ca, connect=connect, pairs_used=0, output_depth=False)
    assert out == read_text(os.path.join(GOLDEN, d, s + ".F.n0D.stdout"))
    assert sv == read_text(os.path.join(GOLDEN, d, s + ".F.n0D.sv"))
    seed = read_text(os.path.join(GOLDEN, d, s + ".sv"))
    sv, out = getsv_oracle.getsv(h, recs, clip, ch, ca, connect=connect, connect_min_mapq=30, seed_text=seed, pairs_used=0, output_depth=False)
    assert out == read_text(os.path.join(GOLDEN, d, s + ".FB.n0D.stdout"))
    assert sv == read_text(os.path.join(GOLDEN, d, s + ".FB.n0D.sv"))
