"""Parity of the CUDA path (through the C ABI / the CLI that sits on it) against the reference's own outputs
(tests/golden/, made by oracle/_ref/seeksv) and against the CPU oracle on seeded inputs. Needs a B200."""
import gzip
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT, read_text

pytestmark = pytest.mark.gpu

GETSV_CASES = [("example", "cancer"), ("example", "normal"), ("micro", "tumor"), ("micro", "normal"),
               ("fuzz", "f11"), ("fuzz", "f12")]   # fuzz: tests/fuzzgen.py through the reference binary (make_golden.py)
CASES = GETSV_CASES + [("kat", "quirks"), ("kat", "start_tid1")]


def _bam(d, s):
    p = os.path.join(GOLDEN, d, s + ".sort.bam")
    return p if os.path.exists(p) else os.path.join(GOLDEN, d, s + ".bam")


def _cli():
    from seeksv_b200 import cli_path
    assert os.path.exists(cli_path()), "build the CLI first (python -m seeksv_b200.build)"
    return cli_path()


def _zcat(path):
    with gzip.open(path, "rb") as f:
        return f.read().decode("latin-1")


@pytest.fixture(scope="module")
def ctx():
    import seeksv_b200
    c = seeksv_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("d,s", CASES)
def test_getclip_cli_bit_exact(d, s, tmp_path):
    pre = str(tmp_path / s)
    r = subprocess.run([_cli(), "getclip", "-o", pre, _bam(d, s)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for ext, name in ((".clip.gz", ".clip.txt"), (".clip.fq.gz", ".clip.fq.txt"),
                      (".unmapped_1.fq.gz", ".unmapped_1.fq.txt"), (".unmapped_2.fq.gz", ".unmapped_2.fq.txt")):
        assert _zcat(pre + ext) == read_text(os.path.join(GOLDEN, d, s + name)), ext


@pytest.mark.parametrize("d,s", CASES)
def test_getclip_c_abi_matches_oracle(ctx, d, s):
    import seeksv_b200
    from oracle import bamio, getclip_oracle
    h, recs = bamio.read_bam(_bam(d, s))
    want = getclip_oracle.getclip(h, recs)
    bam = seeksv_b200.Bam.open(ctx, _bam(d, s))
    assert bam.n_records == len(recs)
    got = bam.getclip()
    for g, w in zip(got, want):
        assert g.decode("latin-1") == w
    # non-default parameters
    want = getclip_oracle.getclip(h, recs, limit=0.8, min_mapq=30, save_low_quality=True)
    got = bam.getclip(match_rate=0.8, min_mapq=30, save_low_quality=True)
    for g, w in zip(got, want):
        assert g.decode("latin-1") == w
    bam.close()


@pytest.mark.parametrize("d,s", GETSV_CASES)
def test_getsv_cli_bit_exact(d, s, tmp_path):
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    out, unm = str(tmp_path / "out.sv"), str(tmp_path / "unm")
    r = subprocess.run([_cli(), "getsv", os.path.join(GOLDEN, d, s + ".clip.sam"), _bam(d, s), clip, out, unm],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".sv"))
    assert r.stdout == read_text(os.path.join(GOLDEN, d, s + ".getsv.stdout"))
    assert os.path.getsize(unm) == 0


@pytest.mark.parametrize("d,normal,tumour", [("example", "normal", "cancer"), ("micro", "normal", "tumor")])
def test_somatic_cli_bit_exact(d, normal, tumour, tmp_path):
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, normal + ".clip.txt")).encode("latin-1"))
    out = str(tmp_path / "somatic.sv")
    r = subprocess.run([_cli(), "somatic", _bam(d, normal), clip, os.path.join(GOLDEN, d, tumour + ".sv"), out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, tumour + ".somatic.temp.sv"))


@pytest.mark.parametrize("d,s", GETSV_CASES)
def test_device_passes_match_oracle(ctx, d, s):
    """insert-size sums, discordant-pair counts and window depth through the C ABI vs the CPU oracle"""
    import random
    import seeksv_b200
    from oracle import bamio, getsv_oracle as G
    h, recs = bamio.read_bam(_bam(d, s))
    bam = seeksv_b200.Bam.open(ctx, _bam(d, s))
    for mq, cap in ((20, 5000000), (0, 1000), (30, 7)):
        n, tot, mean, sq = bam.insert_stats(mq, cap)
        want = G.insert_size_stats(recs, mq, cap)
        if want is None:
            assert n == 0
        else:
            import math
            assert (mean, int(math.sqrt(sq / n))) == want
    mean, dev = G.insert_size_stats(recs, 20, 5000000)
    rng = random.Random(5)
    juncs = []
    for _ in range(300):
        ut, dt = rng.randrange(len(h.names)), rng.randrange(len(h.names))
        if rng.random() < 0.6:
            dt = ut
        up = rng.randrange(1, h.lengths[ut])
        dp = max(1, min(h.lengths[dt], up + rng.randrange(-600, 600))) if dt == ut and rng.random() < 0.7 else rng.randrange(1, h.lengths[dt])
        juncs.append((ut, up, rng.choice("+-"), dt, dp, rng.choice("+-")))
    juncs.append((-1, 5, "+", 0, 5, "+"))
    juncs.append((0, 5, "+", -1, 5, "+"))
    got = bam.discordant_support(juncs, 20, mean, dev, 4)
    for j, g in zip(juncs, got):
        ut, up, us, dt, dp, ds = j
        key = (h.names[ut] if ut >= 0 else "nope", h.names[dt] if dt >= 0 else "nope", us, ds, up, dp)
        assert g == G.discordant_pairs(h, recs, key, 20, mean, dev, 4), j
    # depth over random disjoint windows
    for mq in (20, 0):
        dep = G.depth_arrays(h, recs, mq)
        wins = []
        for tid, ln in enumerate(h.lengths):
            p = 1
            while True:
                p += rng.randrange(1, 700)
                e = p + rng.randrange(0, 500)
                if e > ln:
                    break
                wins.append((tid, p, e))
                p = e + 1
        got = bam.window_depth(wins, mq)
        for (tid, b, e), g in zip(wins, got):
            assert g == [int(x) for x in dep[tid][b:e + 1]], (tid, b, e)
        if mq == 20:
            # the fused call (one stream-ordered sequence, statistics stay on the device) and the rows written by getclip's own pass
            # (with_rows) give the same answers as the three single passes
            want_counts = bam.discordant_support(juncs, 20, mean, dev, 4)
            for fresh in (False, True):
                b2 = bam
                if fresh:
                    b2 = seeksv_b200.Bam.open(ctx, _bam(d, s))
                    b2.getclip_sizes(with_rows=True)
                for cap in (5000000, 1000):
                    st, cnts, deps = b2.getsv_passes(juncs, wins, 20, cap, 4)
                    w2 = G.insert_size_stats(recs, 20, cap)
                    import math
                    assert (st[2], int(math.sqrt(st[3] / st[0]))) == w2
                    assert deps == got
                    if cap == 5000000:
                        assert cnts == want_counts
                    else:
                        assert cnts == b2.discordant_support(juncs, 20, w2[0], w2[1], 4)
                if fresh:
                    b2.close()
    bam.close()


def test_pileup_cap_matches_oracle(ctx, tmp_path):
    """coverage beyond libbam's 8000-read pileup cap (quirk Q12) and the =/X CIGAR quirk"""
    import random
    import seeksv_b200
    from oracle import bamio, getsv_oracle as G
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
                recs.append(bamio.make_rec("r", flag, tid, pos, rng.choice((0, 30, 60)), cig, -1, -1, 0, "A" * 50, "I" * 50))
    path = str(tmp_path / "cap.bam")
    bamio.write_bam(path, h, recs)
    bam = seeksv_b200.Bam.open(ctx, path)
    for mq in (0, 20):
        dep = G.depth_arrays(h, recs, mq)
        wins = [(0, 1, 2500), (0, 2600, 5000), (1, 1, 5000)]
        got = bam.window_depth(wins, mq)
        for (tid, b, e), g in zip(wins, got):
            assert g == [int(x) for x in dep[tid][b:e + 1]], (mq, tid)
    bam.close()


@pytest.mark.parametrize("chunk_log2", [None, "14", "12"])
def test_candidate_order_without_a_sort_matches_oracle(ctx, tmp_path, monkeypatch, chunk_log2):
    """File order of the candidates comes from per-chunk buckets, not from a sort (getclip.cu: cand_group / make_keys): short
    records (hundreds per 16 KiB chunk), most of them clipped on both sides, piled on a few breakpoint keys on two chromosomes,
    so that the greedy clustering sees long runs of equal keys whose members share chunks - any slip in the order inside a
    bucket, or between the '5' and the '3' candidate of one read, changes which read founds a cluster and which CIGAR it keeps.
    A stream this small gets 1 KiB chunks by default; 16 KiB chunks (what a real BAM gets) hold ~300 of these records each."""
    import random
    if chunk_log2:
        monkeypatch.setenv("SEEKSV_B200_CHUNK_LOG2", chunk_log2)
    import seeksv_b200
    from oracle import bamio, getclip_oracle
    rng = random.Random(11)
    h = bamio.Header(["c1", "c2"], [60000, 60000], "@SQ\tSN:c1\tLN:60000\n@SQ\tSN:c2\tLN:60000\n")
    recs = []
    for tid in (0, 1):
        pos = 100
        for _ in range(6000):
            pos += rng.choice((0, 0, 0, 0, 1, 2, 40))
            kind = rng.randrange(6)
            left, right = rng.choice((3, 4, 5, 9)), rng.choice((3, 4, 6))
            mid = rng.choice((8, 9, 10, 12))
            if kind <= 2:
                cig, n = "%dS%dM%dS" % (left, mid, right), left + mid + right
            elif kind == 3:
                cig, n = "%dS%dM" % (left, mid), left + mid
            elif kind == 4:
                cig, n = "%dM%dS" % (mid, right), mid + right
            else:
                cig, n = "%dM" % mid, mid
            seq = "".join(rng.choice("ACGT") if rng.random() < 0.1 else "ACGT"[(pos + j) & 3] for j in range(n))
            qual = "".join(chr(33 + rng.randrange(2, 40)) for _ in range(n))
            recs.append(bamio.make_rec("q%d" % len(recs), rng.choice((0, 16, 0, 1024)), tid, pos, rng.choice((60, 60, 5)), cig, -1, -1, 0, seq, qual))
    path = str(tmp_path / "order.bam")
    bamio.write_bam(path, h, recs)
    want = getclip_oracle.getclip(h, recs)
    bam = seeksv_b200.Bam.open(ctx, path)
    got = bam.getclip()
    for g, w in zip(got, want):
        assert g.decode("latin-1") == w
    bam.close()


def test_no_cpu_fallback():
    """the product never imports the oracle, and the library refuses to run without its CUDA device"""
    import seeksv_b200.lib as lib
    src = open(lib.__file__).read()
    assert "oracle" not in src


@pytest.mark.parametrize("genome,extra", [("chr21:3000000", []), ("chr7:1500000,chr12:1000000,chr3:800000", ["--virus"])])
def test_synthetic_cli_vs_reference_binary(genome, extra, tmp_path):
    """svsim BAM (30x, planted DEL/INV/moved segments[, virus contigs at >5000x]) through both CLIs: every output of
    getclip, getsv and somatic must be byte-identical to the real reference binary (oracle/_ref/seeksv)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "seeksv")
    svsim = os.path.join(ROOT, "seeksv_b200", "bin", "svsim")
    mini = os.path.join(ROOT, "seeksv_b200", "bin", "minialign")
    if not (os.path.exists(ref) and os.path.exists(svsim) and os.path.exists(mini)):
        pytest.skip("needs oracle/_ref/seeksv and the svsim / minialign tools (python bench.py builds them)")
    t = str(tmp_path)
    outs = {}
    for sample in ("tumor", "normal"):
        subprocess.run([svsim, "--out", f"{t}/{sample}", "--genome", genome, "--nsv", "40", "--sample", sample] + extra,
                       check=True, capture_output=True)
    for tag, exe in (("ref", ref), ("b200", _cli())):
        for sample in ("tumor", "normal"):
            pre = f"{t}/{tag}.{sample}"
            bam = f"{t}/{sample}.bam"
            r = subprocess.run([exe, "getclip", "-o", pre, bam], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            if tag == "ref":
                with open(f"{t}/{sample}.clip.sam", "w") as o:
                    subprocess.run([mini, f"{t}/{sample}.fa", pre + ".clip.fq.gz"], check=True, stdout=o)
            r = subprocess.run([exe, "getsv", f"{t}/{sample}.clip.sam", bam, pre + ".clip.gz", pre + ".sv", pre + ".unm"],
                               capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            outs[(tag, sample, "stdout")] = r.stdout
            for ext in (".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz"):
                outs[(tag, sample, ext)] = _zcat(pre + ext)
            outs[(tag, sample, ".sv")] = read_text(pre + ".sv")
        pre = f"{t}/{tag}"
        r = subprocess.run([exe, "somatic", f"{t}/normal.bam", f"{pre}.normal.clip.gz", f"{pre}.tumor.sv", f"{pre}.somatic.sv"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs[(tag, "tumor", "somatic")] = read_text(f"{pre}.somatic.sv")
    n_sv = outs[("ref", "tumor", ".sv")].count("\n")
    assert n_sv > 20, "the planted SVs must be found by the reference itself"
    for (tag, sample, what), v in outs.items():
        if tag == "ref":
            assert outs[("b200", sample, what)] == v, (sample, what)


def _deflate_cases():
    import random
    import zlib
    from oracle import bamio
    rng = random.Random(11)
    raw = bamio.read_bgzf(_bam("micro", "tumor"))
    cases = {"fixture": open(_bam("example", "cancer"), "rb").read()}
    for name, level, strategy, block in (("level0_stored", 0, 0, 0xff00), ("level1", 1, 0, 0xff00), ("level9", 9, 0, 0xff00),
                                         ("fixed", 6, 4, 0xff00), ("tiny_blocks", 6, 0, 777), ("huffman_only", 6, 2, 0x8000),
                                         ("rle", 6, 3, 0xfff0)):
        cases[name] = bamio.bgzf_compress(raw, level, block, strategy)
    # incompressible + highly repetitive payload after a valid BAM (long matches, dist < len copies, long codes, several deflate
    # blocks per BGZF block, a single-code distance alphabet)
    noise = bytes(rng.randrange(256) for _ in range(70000)) + b"A" * 100000 + bytes(range(256)) * 300
    cases["mixed"] = bamio.bgzf_compress(raw + noise, 6)
    # skewed symbol statistics: literal/length codes of up to 15 bits (second-level tables), periodic data that a decoder
    # started at a wrong bit position does not resynchronise on (many speculation rounds)
    skew = bytearray()
    for k in range(60000):
        r = rng.random()
        skew.append(0 if r < 0.5 else 1 if r < 0.75 else 2 if r < 0.87 else rng.randrange(256))
    cases["skewed"] = bamio.bgzf_compress(bytes(skew) + bytes([7, 7, 9]) * 40000 + bytes(rng.randrange(4) for _ in range(50000)), 9)
    return raw, cases


def test_device_inflate_matches_zlib(ctx, tmp_path, monkeypatch):
    """BGZF inflate on the device == zlib, for dynamic / fixed / stored deflate blocks, with the speculative kernel (default),
    the one-decoding-lane kernel (SEEKSV_B200_INFLATE=serial) and the host-thread inflate path (SEEKSV_B200_HOST_INFLATE=1)."""
    import seeksv_b200
    from seeksv_b200.lib import inflate_bgzf
    raw, cases = _deflate_cases()
    for name, img in cases.items():
        want = __import__("gzip").decompress(img)
        for host, kernel in (("0", ""), ("0", "serial"), ("1", "")):
            monkeypatch.setenv("SEEKSV_B200_HOST_INFLATE", host)
            monkeypatch.setenv("SEEKSV_B200_INFLATE", kernel)
            if name in ("mixed", "skewed"):
                # not a well-formed record chain: the first pass that walks it must fail cleanly; the inflate kernel alone
                # must not
                if name == "mixed":
                    bad = seeksv_b200.Bam.from_bgzf(ctx, img)
                    with pytest.raises(seeksv_b200.SvbError):
                        bad.getclip()
                    bad.close()
                assert inflate_bgzf(ctx, img) == want, (name, kernel)
                continue
            bam = seeksv_b200.Bam.from_bgzf(ctx, img)
            assert bam.copy_stream() == want, (name, host, kernel)
            bam.close()


def test_device_inflate_refuses_corrupt_streams(ctx, monkeypatch):
    """A damaged deflate payload must come back as SVB_ERR_FORMAT (or, when the damage happens to decode, as the right number of
    bytes) - never as a fault: the kernel bounds every read by the block's payload and checks LEN against NLEN."""
    import random
    import struct
    import seeksv_b200
    from seeksv_b200.lib import inflate_bgzf
    from oracle import bamio
    raw, cases = _deflate_cases()
    rng = random.Random(5)
    for kernel in ("", "serial"):
        monkeypatch.setenv("SEEKSV_B200_INFLATE", kernel)
        for name in ("level1", "level0_stored", "fixed", "skewed"):
            img = bytearray(cases[name])
            bsize = struct.unpack_from("<H", img, 16)[0] + 1
            for trial in range(6):
                bad = bytearray(img)
                if trial == 0:      # zeroed payload: endless empty stored blocks
                    bad[18:bsize - 8] = bytes(bsize - 26)
                elif trial == 1:    # all ones: reserved block type / invalid codes
                    bad[18:bsize - 8] = b"\xff" * (bsize - 26)
                else:               # a few random bytes damaged inside the first block's payload
                    for _ in range(trial):
                        bad[18 + rng.randrange(bsize - 26)] ^= 1 << rng.randrange(8)
                try:
                    got = inflate_bgzf(ctx, bytes(bad))
                    assert len(got) == len(__import__("gzip").decompress(bytes(img)))
                except seeksv_b200.SvbError:
                    pass
        # the context is still usable
        assert inflate_bgzf(ctx, cases["level1"]) == raw


def test_depth_accounting_closed_form_equals_literal_walk(tmp_path, monkeypatch):
    """getsv with the closed-form range sums vs the literal per-position map walks of bam2depth.cpp:82-124, with flank
    lengths and distances that produce degenerate 0/1-length and wrapped (unsigned) ranges (quirk Q11)."""
    d, s = "micro", "tumor"
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    outs = {}
    for flank in ("200", "1", "0", "30000"):
        for mode in ("closed", "literal"):
            if mode == "literal":
                monkeypatch.setenv("SEEKSV_B200_LITERAL_DEPTH_WALK", "1")
            else:
                monkeypatch.delenv("SEEKSV_B200_LITERAL_DEPTH_WALK", raising=False)
            out = str(tmp_path / ("%s_%s.sv" % (mode, flank)))
            r = subprocess.run([_cli(), "getsv", "-d", "0", "-L", flank, os.path.join(GOLDEN, d, s + ".clip.sam"), _bam(d, s), clip, out,
                                str(tmp_path / "unm")], capture_output=True, text=True, env=dict(os.environ))
            assert r.returncode == 0, r.stderr
            outs[(mode, flank)] = read_text(out) + r.stdout
        assert outs[("closed", flank)] == outs[("literal", flank)], flank


@pytest.mark.parametrize("mode", ["walk10", "walk12"])
def test_alternative_full_pass_forms_give_identical_results(mode, tmp_path, monkeypatch):
    """other walker chunk sizes (more chains, more guesses, other chunk boundaries) produce the same bytes"""
    monkeypatch.setenv("SEEKSV_B200_CHUNK_LOG2", mode[4:])
    for d, s in (("micro", "tumor"), ("example", "cancer"), ("kat", "quirks")):
        pre = str(tmp_path / (mode + s))
        r = subprocess.run([_cli(), "getclip", "-o", pre, _bam(d, s)], capture_output=True, text=True, env=dict(os.environ))
        assert r.returncode == 0, r.stderr
        assert _zcat(pre + ".clip.gz") == read_text(os.path.join(GOLDEN, d, s + ".clip.txt"))
        assert _zcat(pre + ".unmapped_1.fq.gz") == read_text(os.path.join(GOLDEN, d, s + ".unmapped_1.fq.txt"))
        if d == "kat":
            continue
        out = str(tmp_path / (mode + s + ".sv"))
        r = subprocess.run([_cli(), "getsv", os.path.join(GOLDEN, d, s + ".clip.sam"), _bam(d, s), pre + ".clip.gz", out, str(tmp_path / "unm")],
                           capture_output=True, text=True, env=dict(os.environ))
        assert r.returncode == 0, r.stderr
        assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".sv"))


@pytest.mark.parametrize("d,s", [("micro", "tumor"), ("fuzz", "f11")])
def test_run_keeps_the_bam_resident_and_gives_the_same_outputs(d, s, tmp_path):
    """`seeksv run -- getclip ... -- <aligner stand-in> -- getsv ...`: one process, the BAM is loaded once (the second command
    takes the resident copy over), outputs identical to the separate commands / the reference"""
    pre = str(tmp_path / s)
    out = str(tmp_path / (s + ".sv"))
    env = dict(os.environ, SEEKSV_B200_TIMING="1")
    r = subprocess.run([_cli(), "run", "--", "getclip", "-o", pre, _bam(d, s), "--", "test -s %s.clip.fq.gz" % pre, "--",
                        "getsv", os.path.join(GOLDEN, d, s + ".clip.sam"), _bam(d, s), pre + ".clip.gz", out, str(tmp_path / "unm")],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert _zcat(pre + ".clip.gz") == read_text(os.path.join(GOLDEN, d, s + ".clip.txt"))
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".sv"))
    assert r.stdout == read_text(os.path.join(GOLDEN, d, s + ".getsv.stdout"))


@pytest.mark.parametrize("d,s", [("fuzz", "f11"), ("fuzz", "f12"), ("micro", "tumor"), ("example", "cancer")])
def test_chromosome_shards_of_one_bam_reproduce_the_whole_file(ctx, d, s):
    """svb_bam_open_refs cuts an indexed BAM at reference boundaries (.bai virtual offsets); the shards, run one after the
    other as ranks would (prev_tid chain, seeksv_b200/sharding.py), give the whole-file outputs; the getsv passes add up"""
    import seeksv_b200
    from seeksv_b200 import sharding
    from oracle import bamio
    path = _bam(d, s)
    h, recs = bamio.read_bam(path)
    whole = seeksv_b200.Bam.open(ctx, path)
    n_ref = len(h.names)
    golden = [read_text(os.path.join(GOLDEN, d, s + e)) for e in (".clip.txt", ".clip.fq.txt", ".unmapped_1.fq.txt", ".unmapped_2.fq.txt")]
    for world in sorted({1, 2, 3, n_ref, n_ref + 2}):
        workers = [sharding.open_ref_shard(ctx, path, r, world) for r in range(world)]
        assert sum(w.bam.n_records for w in workers) == len(recs)
        for w, (lo, hi) in zip(workers, sharding.assign_chromosomes(h.lengths, world)):
            want_last = [r.tid for r in recs if lo <= r.tid < hi and not r.flag & 12]
            assert w.last_mapped_tid() == (want_last[-1] if want_last else None)
        prev = sharding.prev_tids([w.last_mapped_tid() for w in workers])
        parts = [w.getclip(p) for w, p in zip(workers, prev)]   # (clip, clip.fq, "", "", unmapped-branch records)
        for i in range(2):
            assert "".join(p[i] for p in parts) == golden[i], (world, i)
        # mates are paired by name across the whole file: the merging rank pairs the shards' records
        u1, u2 = workers[0].pair_unmapped(b"".join(p[4] for p in parts))
        assert (u1, u2) == (golden[2], golden[3]), world
        # getsv side: qualifying-pair sums and per-junction counts are additive over shards
        n_w, tot_w, mean_w, _ = whole.insert_stats(20, 5000000)
        stats = [w.bam.insert_stats(20, 5000000) for w in workers]
        assert sum(x[0] for x in stats) == n_w and sum(x[1] for x in stats) == tot_w
        juncs = [(t, p, "+", t, p + 300, "-") for t in range(n_ref) for p in range(100, min(h.lengths[t] - 400, 20000), 1500)]
        want = whole.discordant_support(juncs, 20, mean_w, 25, 4)
        got = [w.bam.discordant_support(juncs, 20, mean_w, 25, 4) for w in workers]
        assert [sum(col) for col in zip(*got)] == list(want)
        for w in workers:
            w.close()
    whole.close()


@pytest.mark.parametrize("d,s", [("fuzz", "f11"), ("fuzz", "f12"), ("micro", "tumor"), ("example", "cancer")])
def test_coordinate_range_shards_of_one_bam_reproduce_the_whole_file(ctx, d, s):
    """range shards inside chromosomes (plan from the .bai's linear index, halo + key ownership on the device, merge per
    chromosome and side, mates paired on the merging rank), run one after the other as ranks would"""
    from seeksv_b200 import sharding
    path = _bam(d, s)
    n_ref = len(seeksv_b200_header(path))
    golden = tuple(read_text(os.path.join(GOLDEN, d, s + e)) for e in (".clip.txt", ".clip.fq.txt", ".unmapped_1.fq.txt", ".unmapped_2.fq.txt"))
    for world in (1, 2, 3, 5, 8):
        plans = sharding.plan_range_shards(path, None, n_ref, world)
        workers = [sharding.RangeShardWorker(ctx, path, p) for p in plans]
        assert all(w.context_has_mapped_record() for w in workers)
        parts = [w.getclip() for w in workers]
        clip, fq = sharding.merge_range_texts([(p[0], p[1]) for p in parts])
        live = [w for w in workers if w.bam is not None]
        u1, u2 = live[0].pair_unmapped(b"".join(p[4] for p in parts))
        assert (clip, fq, u1, u2) == golden, world
        # getsv side on the own regions: record counts, qualifying-pair sums and per-junction counts add up to the whole file's
        import seeksv_b200
        whole = seeksv_b200.Bam.open(ctx, path)
        own = [w.own_view() for w in live]
        assert sum(v.n_records for v in own) == whole.n_records
        n_w, tot_w, mean_w, _ = whole.insert_stats(20, 5000000)
        stats = [v.insert_stats(20, 5000000) for v in own]
        assert sum(x[0] for x in stats) == n_w and sum(x[1] for x in stats) == tot_w
        lens = whole.ref_lens
        juncs = [(t, p, "+", t, p + 300, "-") for t in range(n_ref) for p in range(100, min(lens[t] - 400, 20000), 1500)]
        want = whole.discordant_support(juncs, 20, mean_w, 25, 4)
        got = [v.discordant_support(juncs, 20, mean_w, 25, 4) for v in own]
        assert [sum(col) for col in zip(*got)] == list(want)
        for v in own:
            v.close()
        whole.close()
        for w in workers:
            w.close()


def test_range_shards_refuse_a_read_that_reaches_the_next_shard_from_outside_its_halo(ctx, tmp_path):
    """The documented limit of coordinate-range shards: the halo of a shard comes from the .bai, whose alignment
    end counts M, D, N only, while getclip's '3' breakpoint position also counts `=`. A soft-clipped read with a 60 kb `=` operation
    in front of a cut ends, for getclip, in the NEXT shard's key range, but that shard's halo does not hold the record. The shard
    that has the record must fail loudly (no silent loss of a cluster member); the whole-file run is exact."""
    import random
    import seeksv_b200
    from oracle import bamio, getclip_oracle
    from seeksv_b200 import sharding
    bamtool = os.path.join(ROOT, "oracle", "_ref", "bamtool")
    if not os.path.exists(bamtool):
        pytest.skip("needs oracle/_ref/bamtool (libbam's indexer)")
    rng = random.Random(3)
    h = bamio.Header(["c1"], [200000])
    recs = []
    for i, pos in enumerate(range(100, 150000, 40)):
        recs.append(bamio.make_rec("r%d" % i, 0, 0, pos, 60, "100M", -1, -1, 0, "".join(rng.choice("ACGT") for _ in range(100)), "I" * 100))
    seq = "".join(rng.choice("ACGT") for _ in range(10 + 60000 + 50))
    recs.append(bamio.make_rec("long_eq", 0, 0, 20000, 60, "10M60000=50S", -1, -1, 0, seq, "I" * len(seq)))   # libbam's end 20010, getclip's key 80010
    recs.sort(key=lambda r: (r.tid, r.pos))
    path = str(tmp_path / "halo.bam")
    bamio.write_bam(path, h, recs)
    subprocess.run([bamtool, "index", path], check=True)
    want = getclip_oracle.getclip(h, recs)
    assert want[0].count("\n") == 1 and want[0].split("\t")[1] == "80010"
    whole = seeksv_b200.Bam.open(ctx, path)
    assert tuple(t.decode("latin-1") for t in whole.getclip()) == tuple(want)
    whole.close()
    plans = sharding.plan_range_shards(path, None, 1, 2)
    assert plans[0].key_hi[1] < 80010 and (((plans[0].key_hi[1] - 1) >> 14) - 1) << 14 > 20010, "the cut has to lie between the two ends"
    w0 = sharding.RangeShardWorker(ctx, path, plans[0])
    with pytest.raises(seeksv_b200.SvbError, match="halo"):
        w0.getclip()
    w0.close()
    # (the refusal is conservative: whether the owner of the key happens to hold the record in the context it loads in front of its
    # halo is not something the refusing shard can know)


def seeksv_b200_header(path):
    from oracle import bamio
    return bamio.read_bam(path)[0].names


def test_cli_fails_cleanly_on_damaged_inputs(tmp_path):
    """message on stderr + exit status 1 (the reference's convention), never a crash or a hang"""
    good = open(_bam("micro", "tumor"), "rb").read()
    cases = {
        "truncated.bam": good[:len(good) // 2 + 123],                       # cut in the middle of a BGZF block
        "text.bam": b"this is not a BAM file\n" * 100,
        "empty.bam": b"",
        "payload.bam": good[:5000] + bytes(b ^ 0x5A for b in good[5000:5400]) + good[5400:],   # damaged deflate payload
    }
    for name, data in cases.items():
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        r = subprocess.run([_cli(), "getclip", "-o", str(tmp_path / "o"), p], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1, (name, r.returncode, r.stderr[-300:])
        assert r.stderr.strip(), name
    r = subprocess.run([_cli(), "getclip", "-o", str(tmp_path / "o"), str(tmp_path / "missing.bam")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "fail to open" in r.stderr
    # libbam keeps records inside BGZF blocks, so a file cut at a block boundary is a valid, shorter BAM without the EOF block:
    # the reference reads it (with a warning about the missing EOF marker), and so does this
    p = str(tmp_path / "cut_at_block.bam")
    open(p, "wb").write(good[:_block_boundary(good, len(good) // 2)])
    r = subprocess.run([_cli(), "getclip", "-o", str(tmp_path / "c"), p], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    ref = os.path.join(ROOT, "oracle", "_ref", "seeksv")
    if os.path.exists(ref):
        subprocess.run([ref, "getclip", "-o", str(tmp_path / "cref"), p], capture_output=True, text=True, timeout=120)
        assert _zcat(str(tmp_path / "c.clip.gz")) == _zcat(str(tmp_path / "cref.clip.gz"))
    # header only: a valid BAM without records gives four empty outputs
    from oracle import bamio
    h, _ = bamio.read_bam(_bam("micro", "tumor"))
    p = str(tmp_path / "header_only.bam")
    bamio.write_bam(p, h, [])
    r = subprocess.run([_cli(), "getclip", "-o", str(tmp_path / "h"), p], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert _zcat(str(tmp_path / "h.clip.gz")) == "" and _zcat(str(tmp_path / "h.unmapped_1.fq.gz")) == ""


def _block_boundary(raw, near):
    import struct
    o = 0
    while o < near:
        o += struct.unpack_from("<H", raw, o + 16)[0] + 1
    return o


@pytest.mark.parametrize("d,s", CASES)
def test_device_gzip_images_decompress_to_the_texts(ctx, d, s, tmp_path):
    """svb_getclip with gz_outputs: the four gzip images made on the device (gzip.cu) are valid gzip - Python's zlib checks every
    member's CRC32 and ISIZE - and decompress to exactly the texts of the plain mode; the product's own reader agrees"""
    import seeksv_b200
    import seeksv_b200.lib as L
    bam = seeksv_b200.Bam.open(ctx, _bam(d, s))
    texts = bam.getclip()
    images = bam.getclip_gz()
    bam.close()
    for i, (t, g) in enumerate(zip(texts, images)):
        assert g[:4] == b"\x1f\x8b\x08\x04" and g[12:14] == b"SV", i
        assert gzip.decompress(g) == t, i
        p = str(tmp_path / ("o%d.gz" % i))
        open(p, "wb").write(g)
        assert L.read_gz(p) == t, i
        assert L.read_gz_device(ctx, p) == t, i      # ... and so does the device reader (what getsv reads P.clip.gz with)


def test_device_gzip_on_awkward_texts(ctx, tmp_path):
    """the gzip kernels alone, through svb_gzip_text: empty input, one byte, one symbol, piece and member edges, skewed and
    incompressible bytes (codes longer than 15 bits must be limited)"""
    import random
    import seeksv_b200.lib as L
    rnd = random.Random(11)
    cases = [b"", b"A", b"I" * 300000, b"ACGT" * (16384 * 3) + b"N", rnd.randbytes((1 << 20) + 17), rnd.randbytes(65536),
             bytes(rnd.choices(range(256), weights=[2 ** (-i / 8) for i in range(256)], k=1_500_000)),
             b"".join(bytes([i]) * (2 ** min(i, 21)) for i in range(23))]
    for k, data in enumerate(cases):
        g = L.gzip_text(ctx, data)
        assert gzip.decompress(g) == data, len(data)
        p = str(tmp_path / ("a%d.gz" % k))       # the members hold <= 64 KiB of text each: the BGZF inflate kernel reads them back
        open(p, "wb").write(g)
        assert L.read_gz_device(ctx, p) == data, len(data)


def test_device_gz_reader_refuses_what_it_cannot_read(ctx, tmp_path, monkeypatch):
    """members of 1 MiB of text (the host writer with an explicit zlib level), a foreign gzip file and plain text are SVB_ERR_FORMAT for
    svb_read_gz_device (getsv then reads them on the host); the host writer's default members (64 KiB of text, Huffman only) are read
    on the device like the device writer's; a damaged member of the right form does not take the process down"""
    import seeksv_b200.lib as L
    text = b"chr1\t100\t5\t50M\tACGT\tIIII\tGG\tII\t1\n" * 40000
    host = str(tmp_path / "host.gz")
    L.write_gz(host, text)
    assert L.read_gz_device(ctx, host) == text and L.read_gz(host) == text
    monkeypatch.setenv("SEEKSV_B200_GZ_LEVEL", "6")
    zl = str(tmp_path / "zlib.gz")
    L.write_gz(zl, text)
    monkeypatch.delenv("SEEKSV_B200_GZ_LEVEL")
    foreign = str(tmp_path / "foreign.gz")
    with gzip.open(foreign, "wb") as f:
        f.write(text)
    plain = str(tmp_path / "plain.txt")
    open(plain, "wb").write(text)
    for p in (zl, foreign, plain):
        with pytest.raises(L.SvbError):
            L.read_gz_device(ctx, p)
        assert L.read_gz(p) == text
    good = L.gzip_text(ctx, text)
    bad = bytearray(good)
    for k in range(3000, 3400):
        bad[k] ^= 0x5a
    p = str(tmp_path / "bad.gz")
    open(p, "wb").write(bytes(bad))
    try:
        got = L.read_gz_device(ctx, p)
    except L.SvbError:
        got = None
    # a damaged payload either fails to decode (error) or decodes to other bytes of the declared length (the CRC32 is not checked
    # on this path, include/seeksv_b200.h): never a crash, a hang or a short buffer
    assert got is None or len(got) == len(text)
    ok = str(tmp_path / "ok.gz")
    open(ok, "wb").write(good)
    assert L.read_gz_device(ctx, ok) == text


def test_cli_host_gzip_modes_give_the_same_files(tmp_path):
    d, s = "micro", "tumor"
    want = read_text(os.path.join(GOLDEN, d, s + ".clip.txt"))
    for name, env in (("huff", {"SEEKSV_B200_GZ": "host"}), ("zlib", {"SEEKSV_B200_GZ_LEVEL": "6"})):
        pre = str(tmp_path / name)
        r = subprocess.run([_cli(), "getclip", "-o", pre, _bam(d, s)], capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr
        assert _zcat(pre + ".clip.gz") == want


@pytest.mark.parametrize("d,s", [("micro", "tumor"), ("fuzz", "f11"), ("example", "cancer")])
def test_getsv_seed_file_cli_bit_exact(d, s, tmp_path):
    """getsv -B with the BAM passes on (insert size, discordant pairs, depth of the seeded junctions too)"""
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    out = str(tmp_path / "out.sv")
    r = subprocess.run([_cli(), "getsv", "-B", os.path.join(GOLDEN, d, s + ".sv"), os.path.join(GOLDEN, d, s + ".clip.sam"), _bam(d, s), clip, out,
                        str(tmp_path / "unm")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".B.sv"))
    assert r.stdout == read_text(os.path.join(GOLDEN, d, s + ".B.stdout"))


@pytest.mark.parametrize("by", ["range", "chromosome"])
def test_mgpu_entry_point_single_rank(by, tmp_path):
    """seeksv_b200.mgpu (the torchrun entry point for one BAM on several GPUs) as a single rank: same four files"""
    from seeksv_b200 import mgpu
    d, s = "fuzz", "f12"
    pre = str(tmp_path / by)
    assert mgpu.main(["getclip", "--by", by, "-o", pre, _bam(d, s)]) == 0
    for ext, name in ((".clip.gz", ".clip.txt"), (".clip.fq.gz", ".clip.fq.txt"), (".unmapped_1.fq.gz", ".unmapped_1.fq.txt"),
                      (".unmapped_2.fq.gz", ".unmapped_2.fq.txt")):
        assert _zcat(pre + ext) == read_text(os.path.join(GOLDEN, d, s + name)), ext


@pytest.mark.parametrize("d,s", [("micro", "tumor"), ("fuzz", "f11"), ("example", "cancer")])
def test_mgpu_getsv_single_rank(d, s, tmp_path):
    """seeksv_b200.mgpu getsv as a single rank: the additive shard passes (svb_insert_partial, svb_pairs_depth on a range shard with
    its own-record offset), handed to the host command through SEEKSV_B200_SHARD_RESULTS - same .sv as the single-process command"""
    from seeksv_b200 import mgpu
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    out, unm = str(tmp_path / "out.sv"), str(tmp_path / "unm")
    assert mgpu.main(["getsv", "--", os.path.join(GOLDEN, d, s + ".clip.sam"), _bam(d, s), clip, out, unm]) == 0
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".sv"))


@pytest.mark.parametrize("d,normal,tumour", [("example", "normal", "cancer"), ("micro", "normal", "tumor")])
def test_mgpu_somatic_single_rank(d, normal, tumour, tmp_path):
    from seeksv_b200 import mgpu
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, normal + ".clip.txt")).encode("latin-1"))
    out = str(tmp_path / "somatic.sv")
    assert mgpu.main(["somatic", "--", _bam(d, normal), clip, os.path.join(GOLDEN, d, tumour + ".sv"), out]) == 0
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, tumour + ".somatic.temp.sv"))


@pytest.mark.parametrize("d,s", [("micro", "tumor"), ("fuzz", "f11")])
def test_getsv_connected_reads_cli_bit_exact(d, s, tmp_path):
    """getsv -F with the BAM passes on: discordant pairs and depth of the connected-read junctions as well"""
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    out = str(tmp_path / "out.sv")
    r = subprocess.run([_cli(), "getsv", "-F", os.path.join(GOLDEN, d, s + ".connect.sam"), os.path.join(GOLDEN, d, s + ".clip.sam"), _bam(d, s),
                        clip, out, str(tmp_path / "unm")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == read_text(os.path.join(GOLDEN, d, s + ".F.stdout"))
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".F.sv"))


@pytest.fixture(scope="module")
def c2_prefix(tmp_path_factory):
    """BASELINE.json's C2 workload, generated once per test run: <prefix>.bam / .bam.bai / .fa / .truth.tsv"""
    svsim = os.path.join(ROOT, "seeksv_b200", "bin", "svsim")
    if not os.path.exists(svsim):
        pytest.skip("needs the svsim tool (python -m seeksv_b200.build)")
    pre = str(tmp_path_factory.mktemp("c2") / "c2")
    subprocess.run([svsim, "--out", pre, "--genome", "chr21:46709983", "--cov", "30", "--nsv", "500", "--seed", "20261017"], check=True,
                   stderr=subprocess.DEVNULL)
    return pre


def test_c2_scale_properties(c2_prefix):
    """BASELINE.json's C2 size (chr21-sized chromosome, 30x, ~9.2 M records, 2.7 GB uncompressed), where the Python oracle is out of
    reach: size-independent properties instead - determinism (two runs, identical bytes), shard invariance (8 coordinate-range shards
    and the whole file give the same four outputs), structure of the outputs (positions ascending per chromosome and side, both mate
    files pair up), and the simulator's known insert-size distribution and depth."""
    import hashlib
    import seeksv_b200
    from seeksv_b200 import sharding
    bam = c2_prefix + ".bam"
    ctx = seeksv_b200.Context(0)
    whole = seeksv_b200.Bam.open(ctx, bam)
    n_rec = whole.n_records
    assert 9_000_000 < n_rec < 9_500_000
    first = whole.getclip()
    again = whole.getclip()
    assert [hashlib.md5(t).hexdigest() for t in first] == [hashlib.md5(t).hexdigest() for t in again]
    # shard invariance
    plans = sharding.plan_range_shards(bam, None, 1, 8)
    assert sum(not p.empty for p in plans) == 8
    parts, own_records = [], 0
    for p in plans:
        w = sharding.RangeShardWorker(ctx, bam, p)
        assert w.context_has_mapped_record()
        parts.append(w.getclip())
        v = w.own_view()
        own_records += v.n_records
        v.close()
        if p is plans[-1]:
            u1, u2 = w.pair_unmapped(b"".join(x[4] for x in parts))
        w.close()
    assert own_records == n_rec
    clip, fq = sharding.merge_range_texts([(x[0], x[1]) for x in parts])
    assert (clip.encode("latin-1"), fq.encode("latin-1"), u1.encode("latin-1"), u2.encode("latin-1")) == first
    # structure
    last = {}
    n_lines = 0
    for line in first[0].decode("latin-1").split("\n")[:-1]:
        f = line.split("\t", 3)
        key, pos = (f[0], f[2]), int(f[1])
        assert pos >= last.get(key, 0), line[:80]
        last[key] = pos
        n_lines += 1
    assert first[1].count(b"\n") == 4 * n_lines                      # one FASTQ record per cluster
    assert first[2].count(b"\n") == first[3].count(b"\n") > 0       # the mate files pair up
    # the simulator draws insert sizes from N(500, 25^2) and covers the chromosome 30x with 150 bp reads
    n, tot, mean, sq = whole.insert_stats(20, 5000000)
    assert n > 1_000_000 and 495 <= mean <= 505 and 20 <= int((sq / n) ** 0.5) <= 30
    depth = whole.window_depth([(0, 10_000_000, 10_100_000)], 20)[0]
    assert 24 <= sum(depth) / len(depth) <= 36
    whole.close()
    ctx.close()


@pytest.mark.parametrize("s", ["f106", "e3"])
def test_campaign_fixtures_cli_bit_exact(s, tmp_path):
    """Fixtures from tools/fuzz_campaign.py. f106: a junction at chrA:97 lies in front of the first flank range of the
    smallest-named chromosome, so main_depth's `continue` (bam2depth.cpp:102) leaves its point depth at 0. e3: fuzzgen's edge
    mode - breakpoints and realigned clips inside the first and last 250 bases of every contig (window clamps of the
    discordant-pair and depth passes, wrapped flank ranges) and clipped parts of up to 320 bases. getclip and getsv through the
    CLI against the reference's outputs, closed-form and literal depth accounting."""
    d = "fuzz"
    pre = str(tmp_path / s)
    r = subprocess.run([_cli(), "getclip", "-o", pre, _bam(d, s)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for ext, name in ((".clip.gz", ".clip.txt"), (".clip.fq.gz", ".clip.fq.txt"), (".unmapped_1.fq.gz", ".unmapped_1.fq.txt"),
                      (".unmapped_2.fq.gz", ".unmapped_2.fq.txt")):
        assert _zcat(pre + ext) == read_text(os.path.join(GOLDEN, d, s + name)), ext
    for env in ({}, {"SEEKSV_B200_LITERAL_DEPTH_WALK": "1"}):
        out = str(tmp_path / "out.sv")
        r = subprocess.run([_cli(), "getsv", os.path.join(GOLDEN, d, s + ".clip.sam"), _bam(d, s), pre + ".clip.gz", out,
                            str(tmp_path / "unm")], capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr
        assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".sv")), env
        assert r.stdout == read_text(os.path.join(GOLDEN, d, s + ".getsv.stdout")), env


@pytest.mark.parametrize("s", ["f11", "f12", "f106", "e3"])
def test_somatic_on_fuzz_fixtures_cli_bit_exact(s, tmp_path):
    """somatic of a fuzzed sample against itself (every tumour call finds control support): the host string matching and the
    discordant-pair queries on the 'normal' BAM against the reference's output."""
    d = "fuzz"
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    out = str(tmp_path / "somatic.sv")
    r = subprocess.run([_cli(), "somatic", _bam(d, s), clip, os.path.join(GOLDEN, d, s + ".sv"), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".somatic.temp.sv"))


def _check_digests(pre, want, tmp_path):
    """getclip -> minialign -> getsv (with and without the BAM passes) -> somatic(self) through the CLI on <pre>.bam / <pre>.fa;
    MD5 and size of every output against `want` (a digests.json made by tests/golden/make_c2_digests.py from the reference's run)."""
    import hashlib
    mini = os.path.join(ROOT, "seeksv_b200", "bin", "minialign")
    if not os.path.exists(mini):
        pytest.skip("needs the minialign tool (python -m seeksv_b200.build)")

    def dig(data):
        return {"md5": hashlib.md5(data).hexdigest(), "bytes": len(data)}

    out = str(tmp_path / "b200")
    r = subprocess.run([_cli(), "getclip", "-o", out, pre + ".bam"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for ext in (".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz"):
        with gzip.open(out + ext, "rb") as f:
            assert dig(f.read()) == want[ext], ext
    with open(out + ".clip.sam", "wb") as o:
        subprocess.run([mini, pre + ".fa", out + ".clip.fq.gz"], check=True, stdout=o)
    assert dig(open(out + ".clip.sam", "rb").read()) == want["clip.sam"]
    for tag, extra in (("getsv", []), ("getsv -n 0 -D", ["-n", "0", "-D"])):
        sv = out + (".n0D.sv" if extra else ".sv")
        r = subprocess.run([_cli(), "getsv", *extra, out + ".clip.sam", pre + ".bam", out + ".clip.gz", sv, out + ".unm"],
                           capture_output=True)
        assert r.returncode == 0, r.stderr
        assert dig(open(sv, "rb").read()) == want[tag]["sv"], tag
        assert dig(r.stdout) == want[tag]["stdout"], tag
    r = subprocess.run([_cli(), "somatic", pre + ".bam", out + ".clip.gz", out + ".sv", out + ".somatic"], capture_output=True)
    assert r.returncode == 0, r.stderr
    assert dig(open(out + ".somatic", "rb").read()) == want["somatic (self)"]


def test_c2_full_size_outputs_equal_the_reference_digests(c2_prefix, tmp_path):
    """BASELINE.json's C2 workload at full size (9.2 M records) through the CLI: MD5 and size of every output of getclip, getsv
    (with and without the BAM passes) and somatic against the digests of the reference's own outputs on the same BAM
    (tests/golden/c2/digests.json, made in the build container by tests/golden/make_c2_digests.py - ~2 minutes of single-core
    reference work that the GPU box does not repeat). svsim is deterministic (any thread count), minialign stands in for bwa in
    both arms. Verified on the B200 in round 1 (profiles/r1_late_c2.log)."""
    import json
    with open(os.path.join(GOLDEN, "c2", "digests.json")) as f:
        want = json.load(f)
    assert want["svsim_args"] == ["--genome", "chr21:46709983", "--cov", "30", "--nsv", "500", "--seed", "20261017"]   # = c2_prefix
    _check_digests(c2_prefix, want, tmp_path)


def test_long_clips_cli_bit_exact(tmp_path):
    """tests/golden/long: clipped sequences of 254 / 255 / 256 / 300 / 600 bases. getclip writes them as FASTQ read names, the
    realigner's SAM carries them as QNAME, and libbam's 8-bit l_qname makes every name of 255+ characters unmatchable, which
    shifts the lock-step join (getsv.h:467-505) by one line - getclip, getsv and somatic against the reference's outputs."""
    d, s = "long", "lq"
    pre = str(tmp_path / s)
    r = subprocess.run([_cli(), "getclip", "-o", pre, _bam(d, s)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for ext, name in ((".clip.gz", ".clip.txt"), (".clip.fq.gz", ".clip.fq.txt"), (".unmapped_1.fq.gz", ".unmapped_1.fq.txt"),
                      (".unmapped_2.fq.gz", ".unmapped_2.fq.txt")):
        assert _zcat(pre + ext) == read_text(os.path.join(GOLDEN, d, s + name)), ext
    out = str(tmp_path / "out.sv")
    r = subprocess.run([_cli(), "getsv", os.path.join(GOLDEN, d, s + ".clip.sam"), _bam(d, s), pre + ".clip.gz", out, str(tmp_path / "unm")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".sv"))
    assert r.stdout == read_text(os.path.join(GOLDEN, d, s + ".getsv.stdout"))
    som = str(tmp_path / "somatic.sv")
    r = subprocess.run([_cli(), "somatic", _bam(d, s), pre + ".clip.gz", os.path.join(GOLDEN, d, s + ".sv"), som], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert read_text(som) == read_text(os.path.join(GOLDEN, d, s + ".somatic.temp.sv"))


@pytest.mark.parametrize("d,s", [("micro", "tumor"), ("example", "cancer")])
def test_getclip_reads_sam_text_input(d, s, tmp_path):
    """getclip opens any file whose name does not end in .bam as SAM text (clip_reads.h:367-375). The SAM is libbam's own rendering of
    the fixture BAM (`bamtool bam2sam`); the host converts it back (byte-identical stream, tests/test_sam_text.py) and the device
    passes must give the fixture's outputs."""
    bamtool = os.path.join(ROOT, "oracle", "_ref", "bamtool")
    if not os.path.exists(bamtool):
        pytest.skip("needs oracle/_ref/bamtool")
    sam = str(tmp_path / (s + ".sam"))
    subprocess.run([bamtool, "bam2sam", _bam(d, s), sam], check=True, capture_output=True)
    pre = str(tmp_path / s)
    r = subprocess.run([_cli(), "getclip", "-o", pre, sam], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for ext, name in ((".clip.gz", ".clip.txt"), (".clip.fq.gz", ".clip.fq.txt"), (".unmapped_1.fq.gz", ".unmapped_1.fq.txt"),
                      (".unmapped_2.fq.gz", ".unmapped_2.fq.txt")):
        assert _zcat(pre + ext) == read_text(os.path.join(GOLDEN, d, s + name)), ext


@pytest.fixture(scope="module")
def c3w8_prefix(tmp_path_factory):
    """BASELINE.json's config 3 shape at the C2 size: 24 contigs of 1.95 Mb (9.3 M records), generated once per module"""
    import json
    svsim = os.path.join(ROOT, "seeksv_b200", "bin", "svsim")
    if not os.path.exists(svsim):
        pytest.skip("needs the svsim tool (python -m seeksv_b200.build)")
    with open(os.path.join(GOLDEN, "c2", "c3w8.digests.json")) as f:
        want = json.load(f)
    pre = str(tmp_path_factory.mktemp("c3w8") / "c3w8")
    subprocess.run([svsim, "--out", pre] + want["svsim_args"], check=True, stderr=subprocess.DEVNULL)
    return pre, want


def test_config3_shape_at_c2_size_equals_the_reference_digests(c3w8_prefix, tmp_path):
    """24 chromosomes, 9.3 M records, through the CLI: every output against the reference's digests (tests/golden/c2/c3w8.digests.json)"""
    pre, want = c3w8_prefix
    _check_digests(pre, want, tmp_path)


def test_config3_eight_range_shards_cross_chromosome_boundaries(ctx, c3w8_prefix):
    """The same BAM cut into 8 coordinate-range shards (what 8 ranks would load): every shard spans several of the 1.95 Mb
    chromosomes, so halo, key ownership, the prev_tid hand-over and the per-chromosome merge are all exercised at every cut; the
    shards run one after the other and the merged outputs must have the whole-file digests of the REFERENCE."""
    import hashlib
    from seeksv_b200 import sharding
    pre, want = c3w8_prefix
    path = pre + ".bam"
    whole = __import__("seeksv_b200").Bam.open(ctx, path)
    names, lens, n_records = whole.ref_names, whole.ref_lens, whole.n_records
    n_ref = len(names)
    n_w, tot_w, mean_w, _ = whole.insert_stats(20, 5000000)
    juncs = [(t, p, "+", t, p + 300, "-") for t in range(n_ref) for p in range(1000, lens[t] - 1000, 200000)]
    want_counts = whole.discordant_support(juncs, 20, mean_w, 25, 4)
    whole.close()
    plans = sharding.plan_range_shards(path, None, n_ref, 8)
    assert len([p for p in plans if not p.empty]) == 8
    parts, own_records, stats, got = [], 0, [], []
    for p in plans:      # one shard in HBM at a time
        w = sharding.RangeShardWorker(ctx, path, p)
        assert w.context_has_mapped_record()
        texts = w.bam.getclip(prev_tid=p.prev_tid, export_unmapped=True, key_range=(p.key_lo, p.key_hi), halo_bytes=p.halo_bytes)
        parts.append((texts[0], texts[1], w.bam.last_unmapped_records))
        chroms = {line.split(b"\t", 1)[0] for line in texts[0].split(b"\n")[:-1]}
        assert len(chroms) >= 2, "every shard is meant to cross a chromosome boundary"
        v = w.own_view()
        own_records += v.n_records
        stats.append(v.insert_stats(20, 5000000))
        got.append(v.discordant_support(juncs, 20, mean_w, 25, 4))
        v.close()
        w.close()
    clip, fq = sharding.merge_range_texts_fast([(p[0], p[1]) for p in parts])

    def dig(data):
        return {"md5": hashlib.md5(data).hexdigest(), "bytes": len(data)}
    assert dig(clip) == want[".clip.gz"] and dig(fq) == want[".clip.fq.gz"]
    import seeksv_b200
    mini = seeksv_b200.Bam.from_host(ctx, b"".join(p[2] for p in parts), 0, n_ref)
    mini.set_refs(names, lens)
    cu = mini.getclip_handle(unmapped_only=True)
    assert dig(cu.text(2)) == want[".unmapped_1.fq.gz"] and dig(cu.text(3)) == want[".unmapped_2.fq.gz"]
    cu.close()
    mini.close()
    assert own_records == n_records
    assert sum(x[0] for x in stats) == n_w and sum(x[1] for x in stats) == tot_w
    assert [sum(col) for col in zip(*got)] == list(want_counts)


def test_config4_tumour_normal_pair_at_c2_size_equals_the_reference_digests(tmp_path):
    """BASELINE.json's config 4 shape: a 60x tumour (18.4 M records) and a 30x normal of the same donor, C2-sized chromosome:
    getclip on both, getsv on the tumour, somatic against the normal - every output against the digests of the reference's own run
    (tests/golden/c2/c4.digests.json, tests/golden/make_c2_digests.py <dir> c4)."""
    import hashlib
    import json
    svsim, mini = (os.path.join(ROOT, "seeksv_b200", "bin", t) for t in ("svsim", "minialign"))
    if not (os.path.exists(svsim) and os.path.exists(mini)):
        pytest.skip("needs the svsim / minialign tools (python -m seeksv_b200.build)")
    with open(os.path.join(GOLDEN, "c2", "c4.digests.json")) as f:
        want = json.load(f)

    def dig(data):
        return {"md5": hashlib.md5(data).hexdigest(), "bytes": len(data)}
    pre = {}
    for kind in ("tumor", "normal"):
        pre[kind] = str(tmp_path / kind)
        subprocess.run([svsim, "--out", pre[kind]] + want[kind + "_args"], check=True, stderr=subprocess.DEVNULL)
        r = subprocess.run([_cli(), "getclip", "-o", pre[kind] + ".out", pre[kind] + ".bam"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        for ext in (".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz"):
            with gzip.open(pre[kind] + ".out" + ext, "rb") as f:
                assert dig(f.read()) == want[kind + ext], (kind, ext)
    out = pre["tumor"] + ".out"
    with open(out + ".clip.sam", "wb") as o:
        subprocess.run([mini, pre["tumor"] + ".fa", out + ".clip.fq.gz"], check=True, stdout=o)
    assert dig(open(out + ".clip.sam", "rb").read()) == want["tumor.clip.sam"]
    r = subprocess.run([_cli(), "getsv", out + ".clip.sam", pre["tumor"] + ".bam", out + ".clip.gz", out + ".sv", out + ".unm"], capture_output=True)
    assert r.returncode == 0, r.stderr
    assert dig(open(out + ".sv", "rb").read()) == want["tumor getsv"]["sv"] and dig(r.stdout) == want["tumor getsv"]["stdout"]
    r = subprocess.run([_cli(), "somatic", pre["normal"] + ".bam", pre["normal"] + ".out.clip.gz", out + ".sv", out + ".somatic"], capture_output=True)
    assert r.returncode == 0, r.stderr
    assert dig(open(out + ".somatic", "rb").read()) == want["somatic"]


@pytest.mark.parametrize("name", ["c3mini", "c5mini", "c5"])
def test_other_config_shapes_equal_the_reference_digests(name, tmp_path):
    """Stand-ins for the shapes of BASELINE.json's configs 3 and 5 (24 contigs chr1..chrY whose names sort chr1 < chr10 < ... < chr2;
    a human contig plus HBV / HPV16 contigs at several thousand x), 2.5 M / 1.6 M records, and config 5's shape at full size (c5:
    the C2 chromosome + both virus contigs at 6000x with 40 planted integrations, 9.6 M records): digests of the reference's outputs
    (tests/golden/c2/<name>.digests.json) against the CLI's."""
    import json
    svsim = os.path.join(ROOT, "seeksv_b200", "bin", "svsim")
    if not os.path.exists(svsim):
        pytest.skip("needs the svsim tool (python -m seeksv_b200.build)")
    with open(os.path.join(GOLDEN, "c2", name + ".digests.json")) as f:
        want = json.load(f)
    pre = str(tmp_path / name)
    subprocess.run([svsim, "--out", pre] + want["svsim_args"], check=True, stderr=subprocess.DEVNULL)
    _check_digests(pre, want, tmp_path)
