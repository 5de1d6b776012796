#!/usr/bin/env python
"""Regenerate the golden fixtures under tests/golden/ by running the REAL reference (oracle/_ref/seeksv,
built from /root/reference by oracle/build_ref.sh) plus the reference's bundled bwa 0.7.10.

Run in the build container only (needs /root/reference). The outputs are committed so that the GPU
box - where /root/reference does not exist - can check parity against them.

    python tests/golden/make_golden.py

Layout:
  example/   config C1: the reference's own example BAMs (data fixtures), bwa's clip.sam hand-off,
             and every output of getclip / getsv / somatic (decompressed)
  micro/     seeded simulator output (tests/golden/simulate.py): tumour + normal SAM/BAM with planted
             DEL / INV / DUP / CTX / virus integration, same set of outputs
  kat/       hand-written known-answer BAMs for single quirks; written with
             oracle.bamio.write_bam because the reference's SAM text parser rejects '=' / 'X'
  fuzz/      tests/fuzzgen.py through the whole reference pipeline: f11, f12 (first fixtures), f106 (found by
             tools/fuzz_campaign.py: point depth skipped by bam2depth.cpp:102), e3 (the generator's edge mode: breakpoints at
             the contig ends, clipped parts of up to 320 bases); every sample also as `somatic` against itself
  long/      clipped sequences of 254 - 600 bases (read names of 255+ characters in the realigner's SAM, libbam's 8-bit l_qname)
  c2/        (made by make_c2_digests.py, not by this script) MD5 + size of the reference's outputs on the full-size C2 workload
             and on two smaller stand-ins for the shapes of configs 3 and 5
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SEEKSV_REFERENCE", "/root/reference")
SEEKSV = os.path.join(ROOT, "oracle", "_ref", "seeksv")
BAMTOOL = os.path.join(ROOT, "oracle", "_ref", "bamtool")
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import simulate  # noqa: E402
from oracle import bamio  # noqa: E402


def build_kat(kat):
    """Known-answer records: cluster merge rules, CIGAR op zoo, both-side clips with/without XC,
    filters (dup, mapQ, hard clip), no-quality read, unmapped-mate pairing, chromosome switch."""
    import random
    rng = random.Random(7)

    def rs(n):
        return "".join(rng.choice("ACGT") for _ in range(n))

    def q(n, c="I"):
        return c * n
    ref = rs(1000)
    L = []

    def add(name, flag, tid, pos1, mapq, cigar, seq, qual, mtid=-1, mpos1=0, tlen=0, aux=b""):
        L.append(bamio.make_rec(name, flag, tid, pos1 - 1, mapq, cigar, mtid, mpos1 - 1, tlen, seq, qual, aux))
    clipA = "TTGACCATGA"
    xc = lambda v: b"XCi" + v.to_bytes(4, "little")
    # merge rules: same key, suffix-matching clips of different length, strictly-higher quality wins
    add("m1", 0, 0, 101, 60, "6S30M", clipA[-6:] + ref[100:130], "555555" + q(30))
    add("m2", 0, 0, 101, 60, "10S25M", clipA + ref[100:125], q(10, "H") + q(25, "G"))
    add("m3", 16, 0, 101, 60, "8S35M", "GG" + clipA[-6:] + ref[100:135], "I5IIIIII" + q(35, "J"))
    add("m4", 0, 0, 101, 60, "4S40M", clipA[-4:-1] + "C" + ref[100:140], "IIIK" + q(40))
    add("m5", 0, 0, 101, 60, "10S25M", clipA[:5] + "A" + clipA[6:] + ref[100:112] + "T" + ref[113:125],
        q(10, "J") + q(25, "A"))
    # right clips with I / D / N / = / X (X does not count towards the key position, quirk Q5)
    add("r1", 0, 0, 201, 60, "10M2I10M3D10=1X9M5S",
        ref[200:210] + "AA" + ref[210:220] + ref[223:233] + "T" + ref[234:243] + "CCCCC", q(47))
    add("r2", 0, 0, 206, 60, "20M5N13M6S", ref[205:225] + ref[230:243] + "CCCCCG", q(39))
    add("r3", 0, 0, 211, 60, "33M5S", ref[210:243] + "CCCCC", q(38, "5"))
    # both sides clipped: no XC / XC forward / XC reverse (quirk Q4)
    add("b1", 0, 0, 301, 60, "5S30M7S", "AAAAA" + ref[300:330] + "GGGGGGG", q(42))
    add("b2", 0, 0, 301, 60, "5S30M7S", "AAAAA" + ref[300:330] + "GGGGGGG", q(42), aux=xc(35))
    add("b3", 16, 0, 301, 60, "5S30M7S", "AAAAA" + ref[300:330] + "GGGGGGG", q(42), aux=b"NMC\x00" + xc(35))
    add("x1", 0, 0, 351, 60, "5S30M", "ACGTA" + ref[350:380], q(35), aux=b"XTAU" + xc(30))      # dropped (XC)
    add("x2", 0, 0, 351, 60, "5S30M", "ACGTA" + ref[350:380], q(35), aux=b"XCZabc\x00")         # XC not an int -> 0
    add("x3", 0, 0, 351, 60, "5S30M", "ACGTA" + ref[350:380], q(35), aux=b"XCC\x00")            # XC == 0
    # filters
    add("d1", 1024, 0, 401, 60, "5S30M", "ACGTA" + ref[400:430], q(35))
    add("d2", 0, 0, 401, 0, "5S30M", "ACGTA" + ref[400:430], q(35))
    add("d3", 2048, 0, 401, 60, "5S30M10H", "ACGTA" + ref[400:430], q(35))
    add("d4", 256, 0, 401, 60, "5S30M", "ACGTA" + ref[400:430], q(35))       # secondary is kept
    add("d5", 512, 0, 402, 60, "5S30M", "ACGTA" + ref[401:431], "*")          # QC-fail kept; no qualities
    add("d6", 0, 0, 411, 1, "35M", ref[410:445], q(35))
    add("d7", 0, 0, 412, 60, "3S30M", "nnn".upper() + ref[411:441], q(33))
    # unmapped branch (quirk Q2): mapped read with unmapped mate + the mate; same-end repeat; lone read
    add("u1", 73, 0, 501, 60, "5S30M", "ACGTA" + ref[500:530], q(35), 0, 501)
    add("u1", 133, 0, 501, 0, "*", rs(35), q(35, "F"), 0, 501)
    add("u2", 69, 0, 520, 0, "*", rs(20), q(20), 0, 520)
    add("u2", 69, 0, 521, 0, "*", rs(21), q(21), 0, 520)
    add("u2", 137, 0, 522, 60, "30M", ref[521:551], q(30), 0, 520)
    add("u3", 141, 0, 530, 0, "*", rs(10), q(10))
    add("u4", 77, 0, 531, 0, "*", "", "")
    # chromosome switch (quirk Q1): the first mapped-branch record of chrB is dropped
    add("s1", 0, 1, 101, 60, "5S30M", "ACGTA" + rs(30), q(35))
    add("s2", 0, 1, 101, 60, "5S30M", "ACGTA" + rs(30), q(35))
    add("s3", 0, 1, 101, 60, "5S30M", "ACGTA" + rs(30), q(35))
    add("s5", 0, 1, 150, 60, "31M4S", rs(31) + "ACGT", q(35))
    add("s4", 0, 1, 151, 60, "30M4S", rs(30) + "ACGT", q(34))
    L.sort(key=lambda r: (r.tid, r.pos))
    h = bamio.Header(["chrA", "chrB"], [1000, 1000], "@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:chrA\tLN:1000\n@SQ\tSN:chrB\tLN:1000\n")
    os.makedirs(kat, exist_ok=True)
    bamio.write_bam(os.path.join(kat, "quirks.bam"), h, L)
    # records start on tid 1: the very first record is dropped as well (last_tid starts at 0)
    bamio.write_bam(os.path.join(kat, "start_tid1.bam"), h, [r for r in L if r.tid == 1])


FUZZ_SEEDS, FUZZ_RECORDS = (11, 12, 106), 1800   # 106: found by tools/fuzz_campaign.py (point depth skipped by bam2depth.cpp:102)
FUZZ_EDGE_SEEDS = (3,)   # fuzzgen's edge mode: breakpoints at the contig ends, clipped parts of up to 320 bases
CONNECTED = ("tumor", "f11")          # samples that also get `getsv -F <connected reads>` goldens
SEEDED = ("tumor", "f11", "cancer")   # samples that also get `getsv -B <their own output>` goldens


def build_connect_sam(bam_path, out_path, seed=5, n_reads=160):
    """getsv -F input: "connected read-through reads" - each read is reported as two records with the same name, each soft-clipped
    on one side (process_bwasw.cpp:5-227). Seeded; covers both orientations, both length orders (micro-homology or not),
    repeats of one junction with other lengths, and the records FindJunction skips or leaves pending."""
    import random
    rng = random.Random(seed)
    h, _ = bamio.read_bam(bam_path)
    lines = ["@HD\tVN:1.0\tSO:unsorted"] + ["@SQ\tSN:%s\tLN:%d" % (n, l) for n, l in zip(h.names, h.lengths)]

    def seq(n):
        return "".join(rng.choice("ACGT") for _ in range(n))

    def rec(name, flag, tid, pos1, mapq, cigar, s):
        return "\t".join([name, str(flag), h.names[tid], str(pos1), str(mapq), cigar, "*", "0", "0", s, "I" * len(s)])
    loci = []
    for i in range(n_reads):
        name = "cr%d" % i
        kind = rng.choice(["same", "same", "opp5", "opp3", "bad", "single", "repeat"])
        if kind == "repeat" and loci:
            ta, pa, tb, pb = rng.choice(loci)
            kind = "same"
        else:
            ta, tb = rng.randrange(len(h.names)), rng.randrange(len(h.names))
            pa, pb = rng.randrange(200, h.lengths[ta] - 400), rng.randrange(200, h.lengths[tb] - 400)
            loci.append((ta, pa, tb, pb))
        a_len, b_len = rng.randrange(30, 90), rng.randrange(30, 90)
        total = a_len + b_len - rng.choice([0, 0, 0, 3, 7])           # overlap of the two aligned parts = micro-homology
        s = seq(total)
        mid = "%dM" % a_len if rng.random() < 0.7 else "%dM2I%dM" % (a_len // 2, a_len - a_len // 2 - 2)
        three = rec(name, 0, ta, pa - a_len + 1, rng.choice([0, 1, 30, 60, 60]), mid + "%dS" % (total - a_len), s)       # aligned left part, 3' clipped
        five = rec(name, 0, tb, pb, rng.choice([1, 30, 60, 60]), "%dS%dM" % (total - b_len, b_len), s)                    # aligned right part, 5' clipped
        if kind == "same":
            pair = [three, five] if rng.random() < 0.5 else [five, three]
        elif kind == "opp5":   # both 5' clipped, opposite strands
            pair = [rec(name, 0, ta, pa, 60, "%dS%dM" % (total - a_len, a_len), s), rec(name, 16, tb, pb, 60, "%dS%dM" % (total - b_len, b_len), s)]
        elif kind == "opp3":   # both 3' clipped, opposite strands
            pair = [rec(name, 16, ta, pa, 60, "%dM%dS" % (a_len, total - a_len), s), rec(name, 0, tb, pb, 60, "%dM%dS" % (b_len, total - b_len), s)]
        elif kind == "bad":    # same strand, same side: the second record is ignored and the first stays pending; a third one may fit
            pair = [three, rec(name, 0, tb, pb, 60, "%dM%dS" % (b_len, total - b_len), s), five]
        else:
            pair = [rng.choice([three, five])]
        lines.extend(pair)
        if rng.random() < 0.1:   # records FindJunction skips: duplicate, hard clip, clipped on both sides, not clipped, unmapped
            lines.append(rec("skip%d" % i, 1024, ta, pa, 60, "%dM%dS" % (a_len, total - a_len), s))
            lines.append(rec("skip%d" % i, 0, ta, pa, 60, "5H%dM%dS" % (a_len, total - a_len - 5), s[5:]))
            lines.append(rec("skip%d" % i, 0, ta, pa, 60, "5S%dM%dS" % (a_len - 5, total - a_len), s))
            lines.append(rec("skip%d" % i, 0, ta, pa, 60, "%dM" % total, s))
            lines.append(rec("skip%d" % i, 4, ta, pa, 60, "%dM%dS" % (a_len, total - a_len), s))
    with open(out_path, "w") as f:
        f.write("\n".join(lines) + "\n")


def build_long_clips(path):
    import random
    rng = random.Random(3)
    g = "".join(rng.choice("ACGT") for _ in range(20000))
    recs = []
    for j, (p, src, k_max) in enumerate([(5000, 12000, 300), (7000, 15000, 254), (9000, 1000, 255), (11000, 3000, 256),
                                         (13000, 17000, 600), (14000, 2000, 120)]):
        for r in range(4):      # left clips whose lengths differ by one merge into one cluster (suffix match)
            k = k_max - r
            seq = g[src + k_max - k:src + k_max] + g[p:p + 100]
            recs.append((p, bamio.make_rec("b%d_%d" % (j, r), 0, 0, p, 60, "%dS100M" % k, -1, -1, 0, seq, "I" * len(seq), b"")))
        for r in range(3):      # right clips
            k = k_max - r
            seq = g[p + 200:p + 300] + g[src:src + k]
            recs.append((p + 200, bamio.make_rec("c%d_%d" % (j, r), 0, 0, p + 200, 60, "100M%dS" % k, -1, -1, 0, seq, "I" * len(seq), b"")))
    recs.sort(key=lambda x: x[0])
    hdr = bamio.Header(["chr1"], [20000], "@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:chr1\tLN:20000\n")
    bamio.write_bam(path, hdr, [r for _, r in recs])
    return g


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, **kw)


def gunzip_to(src, dst):
    with gzip.open(src, "rb") as f, open(dst, "wb") as o:
        o.write(f.read())


def strip_pg(sam_in, sam_out):
    with open(sam_in) as f, open(sam_out, "w") as o:
        for line in f:
            if not line.startswith("@PG"):
                o.write(line)


def pipeline(work, outdir, bwa, fasta, samples, somatic_pair=None, getsv_args=()):
    """getclip -> bwa mem -> getsv for each sample, then somatic(normal, tumour)."""
    for s in samples:
        bam = os.path.join(outdir, s + ".sort.bam")
        pre = os.path.join(work, s)
        run([SEEKSV, "getclip", "-o", pre, bam], stderr=subprocess.DEVNULL)
        for ext, name in ((".clip.gz", ".clip.txt"), (".clip.fq.gz", ".clip.fq.txt"),
                          (".unmapped_1.fq.gz", ".unmapped_1.fq.txt"), (".unmapped_2.fq.gz", ".unmapped_2.fq.txt")):
            gunzip_to(pre + ext, os.path.join(outdir, s + name))
        with open(pre + ".clip.raw.sam", "w") as o:
            run([bwa, "mem", fasta, pre + ".clip.fq.gz"], stdout=o, stderr=subprocess.DEVNULL)
        strip_pg(pre + ".clip.raw.sam", os.path.join(outdir, s + ".clip.sam"))
        with open(os.path.join(outdir, s + ".getsv.stdout"), "w") as o:
            run([SEEKSV, "getsv", *getsv_args, os.path.join(outdir, s + ".clip.sam"), bam, pre + ".clip.gz",
                 os.path.join(outdir, s + ".sv"), pre + ".clipunmap"], stdout=o, stderr=subprocess.DEVNULL)
        assert os.path.getsize(pre + ".clipunmap") == 0     # quirk Q7
        # host-only mode (no insert size, no discordant pairs, no depth): exercises join / merge / filters alone
        with open(os.path.join(outdir, s + ".n0D.stdout"), "w") as o:
            run([SEEKSV, "getsv", "-n", "0", "-D", os.path.join(outdir, s + ".clip.sam"), bam, pre + ".clip.gz",
                 os.path.join(outdir, s + ".n0D.sv"), pre + ".clipunmap"], stdout=o, stderr=subprocess.DEVNULL)
        # -F: junctions from connected read-through reads (FindJunction, process_bwasw.cpp:5-227), alone and together with -B
        if s in CONNECTED:
            connect = os.path.join(outdir, s + ".connect.sam")
            build_connect_sam(bam, connect)
            for tag, extra in ((".F", ("-F", connect)), (".F.n0D", ("-F", connect, "-n", "0", "-D")),
                               (".FB.n0D", ("-F", connect, "-B", os.path.join(outdir, s + ".sv"), "-w", "30", "-n", "0", "-D"))):
                with open(os.path.join(outdir, s + tag + ".stdout"), "w") as o:
                    run([SEEKSV, "getsv", *extra, os.path.join(outdir, s + ".clip.sam"), bam, pre + ".clip.gz",
                         os.path.join(outdir, s + tag + ".sv"), pre + ".clipunmap"], stdout=o, stderr=subprocess.DEVNULL)
        # -B: the junctions of an earlier output seed the map (ReadBreakpoint, getsv.cpp:1291-1323); with and without the BAM passes
        if s in SEEDED:
            for tag, extra in ((".B", ()), (".B.n0D", ("-n", "0", "-D"))):
                with open(os.path.join(outdir, s + tag + ".stdout"), "w") as o:
                    run([SEEKSV, "getsv", "-B", os.path.join(outdir, s + ".sv"), *extra, os.path.join(outdir, s + ".clip.sam"), bam,
                         pre + ".clip.gz", os.path.join(outdir, s + tag + ".sv"), pre + ".clipunmap"], stdout=o, stderr=subprocess.DEVNULL)
    if somatic_pair:
        normal, tumour = somatic_pair
        run([SEEKSV, "somatic", os.path.join(outdir, normal + ".sort.bam"), os.path.join(work, normal + ".clip.gz"),
             os.path.join(outdir, tumour + ".sv"), os.path.join(outdir, tumour + ".somatic.temp.sv")],
            stderr=subprocess.DEVNULL)


def main():
    assert os.path.isdir(REF), "needs the reference checkout"
    run([os.path.join(ROOT, "oracle", "build_ref.sh")])
    work = tempfile.mkdtemp(prefix="golden_")
    bwa = os.path.join(work, "bwa")
    shutil.copy(os.path.join(REF, "example", "bin", "bwa"), bwa)
    os.chmod(bwa, 0o755)

    # ---- C1: bundled example ------------------------------------------------------------------
    ex = os.path.join(HERE, "example")
    os.makedirs(ex, exist_ok=True)
    for s in ("cancer", "normal"):
        for ext in (".sort.bam", ".sort.bam.bai"):
            shutil.copy(os.path.join(REF, "example", s + ext), os.path.join(ex, s + ext))
            os.chmod(os.path.join(ex, s + ext), 0o644)
    refdir = os.path.join(work, "exref")
    shutil.copytree(os.path.join(REF, "example", "reference"), refdir)
    pipeline(work, ex, bwa, os.path.join(refdir, "example.fa"), ("normal", "cancer"), ("normal", "cancer"))

    # ---- micro: simulated tumour / normal with every SV class -----------------------------------
    mi = os.path.join(HERE, "micro")
    os.makedirs(mi, exist_ok=True)
    genome = None
    for kind in ("tumor", "normal"):
        genome, header, lines = simulate.micro_sample(kind)
        sam = os.path.join(work, kind + ".sam")
        with open(sam, "w") as f:
            f.write("\n".join(header + lines) + "\n")
        run([BAMTOOL, "sam2bam", sam, os.path.join(mi, kind + ".sort.bam")], stderr=subprocess.DEVNULL)
        run([BAMTOOL, "index", os.path.join(mi, kind + ".sort.bam")])
    fa = os.path.join(work, "micro.fa")
    simulate.write_fasta(genome, fa)
    run([bwa, "index", fa], stderr=subprocess.DEVNULL)
    # -n 100000: keep the insert-size / discordant-pair step on although the sample is small
    pipeline(work, mi, bwa, fa, ("normal", "tumor"), ("normal", "tumor"))

    # ---- known-answer SAMs ------------------------------------------------------------------------
    kat = os.path.join(HERE, "kat")
    build_kat(kat)
    for name in sorted(os.listdir(kat)):
        if not name.endswith(".bam"):
            continue
        pre = os.path.join(work, name[:-4])
        run([SEEKSV, "getclip", "-o", pre, os.path.join(kat, name)], stderr=subprocess.DEVNULL)
        for ext, out in ((".clip.gz", ".clip.txt"), (".clip.fq.gz", ".clip.fq.txt"),
                         (".unmapped_1.fq.gz", ".unmapped_1.fq.txt"), (".unmapped_2.fq.gz", ".unmapped_2.fq.txt")):
            gunzip_to(pre + ext, os.path.join(kat, name[:-4] + out))
    # ---- fuzz: awkward records (tests/fuzzgen.py), the full pipeline of the reference on them -------------------
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fuzzgen
    fz = os.path.join(HERE, "fuzz")
    os.makedirs(fz, exist_ok=True)
    for seed, edge in [(x, False) for x in FUZZ_SEEDS] + [(x, True) for x in FUZZ_EDGE_SEEDS]:
        name = ("e%d" if edge else "f%d") % seed
        _, _, genome = fuzzgen.write(os.path.join(fz, name + ".sort.bam"), seed, FUZZ_RECORDS, edge)
        run([BAMTOOL, "index", os.path.join(fz, name + ".sort.bam")])
        fa = os.path.join(work, name + ".fa")
        fuzzgen.write_fasta(genome, fa)
        run([bwa, "index", fa], stderr=subprocess.DEVNULL)
        pipeline(work, fz, bwa, fa, (name,), (name, name))     # somatic against itself: every call has control support
    # ---- long clips: clipped sequences of 254 / 255 / 256 / 300 / 600 bases become read names of clip.sam; libbam keeps
    # l_qname in 8 bits, so the names of 255+ characters never match their clip line again and the lock-step join of
    # getsv.h:467-505 pairs the alignment with the NEXT line (found by probing, tests/test_sam_text.py)
    lg = os.path.join(HERE, "long")
    os.makedirs(lg, exist_ok=True)
    genome = build_long_clips(os.path.join(lg, "lq.sort.bam"))
    run([BAMTOOL, "index", os.path.join(lg, "lq.sort.bam")])
    fa = os.path.join(work, "lq.fa")
    with open(fa, "w") as f:
        f.write(">chr1\n" + genome + "\n")
    run([bwa, "index", fa], stderr=subprocess.DEVNULL)
    pipeline(work, lg, bwa, fa, ("lq",), ("lq", "lq"))
    shutil.rmtree(work)
    print("golden fixtures regenerated")


if __name__ == "__main__":
    main()
