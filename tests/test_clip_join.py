"""clip_join: the device join of P.clip.gz lines with the realigned clip alignments.

Three layers, each checked against the one below it:
  * the host mirror of the reference's loop (host/junction.cpp: join_clips_with_alignments) is pinned by the goldens of the
    reference binary (test_gpu_parity.py, test_campaign.py: .sv outputs byte for byte);
  * the device join's RULES (csrc/clipjoin_core.h, one source for host and device) run on the CPU in tools/clipjoin_sim.cpp and
    must build the same junction map as the host mirror - here, without a GPU, on every fixture and on mutated fixtures
    (alignments dropped / doubled / swapped / hard-clipped / truncated, lines dropped / doubled / swapped: the cases in which
    runs of lines and blocks of alignments fall out of step), also with every chunk's entry guess deliberately wrong;
  * the KERNELS (csrc/clipjoin.cu through svb_clip_join) must return exactly the candidates of that CPU run (-m gpu), and
    `seeksv getsv` - which joins on the device unless SEEKSV_B200_DEVICE_JOIN=0 - must write the reference's outputs.
"""
import glob
import gzip
import os
import random
import subprocess

import pytest

from conftest import GOLDEN, ROOT, read_text

SIM = os.path.join(ROOT, "seeksv_b200", "bin", "clipjoin_sim")
FIXTURES = sorted(p[:-len(".clip.sam")] for p in glob.glob(os.path.join(GOLDEN, "*", "*.clip.sam")))


@pytest.fixture(scope="module")
def sim():
    from seeksv_b200 import build
    build.build_tools()
    assert os.path.exists(SIM)
    return SIM


def _mutate(prefix, rng, out_dir):
    """a fixture's clip.sam / clip.txt with a few random edits that break the lock step of lines and alignments"""
    sam = read_text(prefix + ".clip.sam").split("\n")
    hdr = [l for l in sam if l.startswith("@")]
    body = [l for l in sam if l and not l.startswith("@")]
    clip = [l for l in read_text(prefix + ".clip.txt").split("\n") if l]
    for _ in range(rng.randrange(1, 12)):
        op = rng.randrange(9)
        if op == 0 and body:
            del body[rng.randrange(len(body))]
        elif op == 1 and body:
            i = rng.randrange(len(body))
            body.insert(i, body[i])
        elif op == 2 and len(body) > 2:
            i = rng.randrange(len(body) - 1)
            body[i], body[i + 1] = body[i + 1], body[i]
        elif op == 3 and clip:
            del clip[rng.randrange(len(clip))]
        elif op == 4 and clip:
            i = rng.randrange(len(clip))
            clip.insert(i, clip[i])
        elif op == 5 and body:
            i = rng.randrange(len(body))
            f = body[i].split("\t")
            if f[5] != "*":
                f[5] = "5H" + f[5] if rng.random() < .5 else f[5] + "7H"
            body[i] = "\t".join(f)
        elif op == 6 and body:
            i = rng.randrange(len(body))
            f = body[i].split("\t")
            f[1] = str(int(f[1]) ^ rng.choice((4, 16, 256)))
            body[i] = "\t".join(f)
        elif op == 7 and body:
            body = body[:rng.randrange(len(body))]          # the alignments run out
        elif op == 8 and len(clip) > 3:
            i = rng.randrange(len(clip) - 1)
            clip[i], clip[i + 1] = clip[i + 1], clip[i]
    sam_path, clip_path = os.path.join(out_dir, "m.clip.sam"), os.path.join(out_dir, "m.clip.txt")
    with open(sam_path, "w", encoding="latin-1", newline="") as f:
        f.write("\n".join(hdr + body) + "\n")
    with open(clip_path, "w", encoding="latin-1", newline="") as f:
        f.write("\n".join(clip) + "\n")
    return sam_path, clip_path


def _run_sim(sim, sam, clip, *extra):
    r = subprocess.run([sim, sam, clip] + list(extra), capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), (sam, extra, r.stdout, r.stderr)
    return [int(x) for x in r.stdout.split()[1:]]


@pytest.mark.parametrize("prefix", FIXTURES, ids=[os.path.relpath(p, GOLDEN) for p in FIXTURES])
def test_rules_build_the_host_mirrors_map(sim, prefix):
    runs, cands, entries, rounds = _run_sim(sim, prefix + ".clip.sam", prefix + ".clip.txt")
    assert rounds == 0, "well-formed input: every entry guess is right"
    again = _run_sim(sim, prefix + ".clip.sam", prefix + ".clip.txt", "--wrong-guesses")
    assert again[:3] == [runs, cands, entries]


def test_rules_on_inputs_that_fall_out_of_step(sim, tmp_path):
    rng = random.Random(20261018)
    bases = [p for p in FIXTURES if os.path.basename(p) in ("f11", "e3", "tumor", "cancer", "lq")]
    for _ in range(120):
        sam, clip = _mutate(rng.choice(bases), rng, str(tmp_path))
        _run_sim(sim, sam, clip)
        _run_sim(sim, sam, clip, "--wrong-guesses")


def _device_equals_cpu(ctx, sim, sam, clip, work):
    from seeksv_b200 import lib
    os.makedirs(work, exist_ok=True)
    _run_sim(sim, sam, clip, "--dump", work)
    arr = {}
    for name in ("lines", "seqs", "alns", "names", "cigars", "cands"):
        with open(os.path.join(work, name + ".bin"), "rb") as f:
            arr[name] = f.read()
    got = lib.clip_join_raw(ctx, arr["lines"], arr["seqs"], arr["alns"], arr["names"], arr["cigars"])
    assert got == arr["cands"], (sam, len(got) // lib.JOIN_CAND_BYTES, len(arr["cands"]) // lib.JOIN_CAND_BYTES)
    return len(got) // lib.JOIN_CAND_BYTES


@pytest.fixture(scope="module")
def ctx():
    import seeksv_b200
    c = seeksv_b200.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("prefix", FIXTURES, ids=[os.path.relpath(p, GOLDEN) for p in FIXTURES])
def test_device_candidates_equal_the_cpu_run_of_the_rules(ctx, sim, prefix, tmp_path):
    _device_equals_cpu(ctx, sim, prefix + ".clip.sam", prefix + ".clip.txt", str(tmp_path / "d"))


@pytest.mark.gpu
def test_device_candidates_on_inputs_that_fall_out_of_step(ctx, sim, tmp_path):
    """mutated fixtures: wrong entry guesses on the device (repair rounds), exhausted alignment streams, hard clips, duplicates"""
    rng = random.Random(7)
    bases = [p for p in FIXTURES if os.path.basename(p) in ("f11", "e3", "tumor", "cancer", "lq")]
    total = 0
    for i in range(40):
        sam, clip = _mutate(rng.choice(bases), rng, str(tmp_path))
        total += _device_equals_cpu(ctx, sim, sam, clip, str(tmp_path / "d"))
    assert total > 0


@pytest.mark.gpu
def test_device_join_of_a_large_synthetic_input(ctx, sim, tmp_path):
    """tens of thousands of runs (hundreds of chunks, several tiles per scan / sort): lines and alignments generated in step, then
    a few hundred random edits so that some chunks start from wrong guesses"""
    rng = random.Random(3)
    chroms = ["c%d" % i for i in range(1, 6)]
    hdr = ["@SQ\tSN:%s\tLN:1000000" % c for c in chroms]
    lines, body = [], []
    for i in range(30000):
        seq = "".join(rng.choice("ACGT") for _ in range(rng.randrange(12, 40)))
        reps = 1 if rng.random() < 0.8 else rng.randrange(2, 4)
        for _ in range(reps):
            side = rng.choice("53")
            lines.append("\t".join([rng.choice(chroms), str(rng.randrange(1, 900000)), side, "30M", "A" * 30, "I" * 30, seq, "I" * len(seq), str(rng.randrange(1, 9))]))
        for a in range(rng.choice((1, 1, 1, 2, 3))):
            unmapped = rng.random() < 0.3
            flag = 4 if unmapped else rng.choice((0, 16, 256, 272, 0, 16))
            cig = "*" if unmapped else rng.choice(("%dM" % len(seq), "5S%dM" % (len(seq) - 5), "%dM3H" % len(seq), "4H%dM" % len(seq), "%dM2D3M" % (len(seq) - 3)))
            body.append("\t".join([seq, str(flag), "*" if unmapped else rng.choice(chroms), "0" if unmapped else str(rng.randrange(1, 900000)),
                                   str(rng.choice((0, 30, 60))), cig, "*", "0", "0", seq, "*"]))
    for _ in range(300):
        op = rng.randrange(4)
        if op == 0:
            del body[rng.randrange(len(body))]
        elif op == 1:
            i = rng.randrange(len(body))
            body.insert(i, body[i])
        elif op == 2:
            del lines[rng.randrange(len(lines))]
        else:
            i = rng.randrange(len(body) - 1)
            body[i], body[i + 1] = body[i + 1], body[i]
    sam, clip = str(tmp_path / "big.clip.sam"), str(tmp_path / "big.clip.txt")
    with open(sam, "w") as f:
        f.write("\n".join(hdr + body) + "\n")
    with open(clip, "w") as f:
        f.write("\n".join(lines) + "\n")
    n = _device_equals_cpu(ctx, sim, sam, clip, str(tmp_path / "d"))
    assert n > 10000


@pytest.mark.gpu
@pytest.mark.parametrize("join", ["1", "0"])
@pytest.mark.parametrize("d,s", [("example", "cancer"), ("micro", "tumor"), ("fuzz", "f11"), ("fuzz", "f12"), ("long", "lq")])
def test_getsv_cli_with_either_join_bit_exact(d, s, join, tmp_path):
    """SEEKSV_B200_DEVICE_JOIN=1 (the default of `getsv` when it has BAM passes to run) and =0 (the host mirror)"""
    from seeksv_b200 import cli_path
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    bam = os.path.join(GOLDEN, d, s + ".sort.bam")
    if not os.path.exists(bam):
        bam = os.path.join(GOLDEN, d, s + ".bam")
    out, unm = str(tmp_path / "out.sv"), str(tmp_path / "unm")
    env = dict(os.environ, SEEKSV_B200_DEVICE_JOIN=join)
    r = subprocess.run([cli_path(), "getsv", os.path.join(GOLDEN, d, s + ".clip.sam"), bam, clip, out, unm], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".sv"))
    assert r.stdout == read_text(os.path.join(GOLDEN, d, s + ".getsv.stdout"))
