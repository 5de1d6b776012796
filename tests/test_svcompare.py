"""tools/svcompare.cpp (seeksv_b200/bin/svcompare) against the reference's stand-alone evaluator built from its own source
(oracle/_ref/svcompare, oracle/build_ref.sh): output file, stdout and exit status on the golden call files and on seeded random
call sets whose positions crowd inside the +-50 bp window. CPU only; SURVEY.md section 8(f) item 4."""
import os
import random
import subprocess

import pytest

from conftest import GOLDEN, ROOT

REF = os.path.join(ROOT, "oracle", "_ref", "svcompare")
OURS = os.path.join(ROOT, "seeksv_b200", "bin", "svcompare")
pytestmark = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(OURS)),
                                reason="needs oracle/_ref/svcompare and seeksv_b200/bin/svcompare (python -m seeksv_b200.build)")


def _both(args, tmp_path):
    res = []
    for tag, exe in (("ref", REF), ("ours", OURS)):
        out = str(tmp_path / (tag + ".out"))
        r = subprocess.run([exe] + args + [out], capture_output=True, text=True, cwd=str(tmp_path))
        res.append((r.returncode, r.stdout, open(out).read() if os.path.exists(out) else None))
    assert res[0] == res[1]
    return res[0]


@pytest.mark.parametrize("control,target", [("example/normal.sv", "example/cancer.sv"), ("example/cancer.sv", "example/cancer.sv"),
                                            ("fuzz/f11.sv", "fuzz/f12.sv"), ("micro/normal.sv", "micro/tumor.sv"),
                                            ("fuzz/f11.n0D.sv", "fuzz/f11.sv"), ("long/lq.sv", "long/lq.n0D.sv")])
def test_seeksv_mode_on_golden_call_files(control, target, tmp_path):
    rc, stdout, out = _both(["seeksv", os.path.join(GOLDEN, control), os.path.join(GOLDEN, target)], tmp_path)
    assert rc == 0 and out is not None and len(stdout.split()) == 2


def _anchors(rng, n):
    chrs = ["chr2", "chr10", "chrX"]
    return [(rng.choice(chrs), rng.randrange(1000, 5000), rng.choice("+-"), rng.choice(chrs), rng.randrange(1000, 5000), rng.choice("+-"))
            for _ in range(n)]


def _random_calls(rng, n, crest, anchors):
    lines = [] if crest else ["@left_chr\tleft_pos\tleft_strand\tleft_clip_read_NO\tright_chr\tright_pos\tright_strand\tright_clip_read_NO\t"
                              "microhomology_length\tabnormal_readpair_NO\tsvtype\tmore"]
    for _ in range(n):
        uc, up, us, dc, dp, ds = rng.choice(anchors)
        up += rng.choice([0, 0, 1, -1, 49, 50, 51, -50, -51, 100, rng.randrange(-120, 120)])
        dp += rng.choice([0, 0, 1, -1, 49, 50, 51, -50, -51, 100, rng.randrange(-120, 120)])
        if rng.random() < 0.2:
            us, ds = rng.choice("+-"), rng.choice("+-")
        a, b = rng.randrange(0, 30), rng.randrange(0, 30)
        ty = rng.choice(["DEL", "INS", "INV", "CTX", "ITX"])
        if crest:
            lines.append("\t".join(map(str, [uc, up, us, a, dc, dp, ds, b, ty, "x", "y"])))
        else:
            lines.append("\t".join(map(str, [uc, up, us, a, dc, dp, ds, b, rng.randrange(-1, 9), rng.randrange(0, 5), ty, "tail", "more"])))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", range(12))
def test_random_call_sets_all_modes(seed, tmp_path):
    rng = random.Random(seed)
    mode = ["seeksv", "crest"][seed & 1]
    crest_target = bool(seed & 2)
    anchors = _anchors(rng, 8)      # both files crowd around the same junctions
    control = _random_calls(rng, rng.randrange(0, 60), mode == "crest", anchors)
    target = _random_calls(rng, rng.randrange(1, 60), crest_target, anchors)
    if seed % 3 == 0:      # make sure exact and near matches exist
        target += "".join(l + "\n" for l in control.splitlines()[1::3]) if (mode == "crest") == crest_target else ""
    (tmp_path / "control.txt").write_text(control)
    (tmp_path / "target.txt").write_text(target)
    rc, stdout, out = _both([mode] + (["-t"] if crest_target else []) + ["control.txt", "target.txt"], tmp_path)
    assert rc == 0


@pytest.mark.parametrize("seed", range(8))
def test_simu_mode(seed, tmp_path):
    rng = random.Random(100 + seed)
    sv, cnv = [], []
    for _ in range(rng.randrange(1, 12)):
        sv.append("%s\t%d\t%d\tA\tx\t." % (rng.choice(["inv", "INV", "tra"]), rng.randrange(500, 9000), rng.randrange(50, 800)))
    for _ in range(rng.randrange(1, 12)):
        s = rng.randrange(500, 9000)
        e = s + rng.randrange(50, 800)
        if rng.random() < 0.5:
            cnv.append("ldel\t%d\t%d\tA\t." % (s, e))
        else:
            cnv.append("lins\t%d\t%d\tA\t.\t%s" % (s, e, ";".join("A:%d" % rng.randrange(500, 9000) for _ in range(rng.randrange(1, 4)))))
    calls = ["@left_chr\tx"]
    for line in sv + cnv:        # calls near the planted events
        f = line.split("\t")
        p, q = int(f[1]), int(f[2])
        if f[0].lower() == "inv":
            q = p + q - 1
            calls.append("\t".join(map(str, ["chr17", p - 1 + rng.randrange(-60, 60), "+", 5, "chr17", q + rng.randrange(-60, 60), "-", 4, 0, 1, "INV", "t"])))
        elif f[0] == "ldel":
            calls.append("\t".join(map(str, ["chr17", p - 1 + rng.randrange(-60, 60), "+", 5, "chr17", q + 1 + rng.randrange(-60, 60), "+", 4, 0, 1, "DEL", "t"])))
    (tmp_path / "sv.txt").write_text("\n".join(sv) + "\n")
    (tmp_path / "cnv.txt").write_text("\n".join(cnv) + "\n")
    (tmp_path / "calls.txt").write_text("\n".join(calls) + "\n")
    (tmp_path / "narea.txt").write_text("".join("chr17\t%d\t%d\n" % (b, b + rng.randrange(10, 400)) for b in sorted(rng.sample(range(400, 9000), 3))))
    args = ["simu", "-c", "chr17"] + (["-n", "narea.txt"] if seed & 1 else []) + ["sv.txt", "cnv.txt", "calls.txt"]
    rc, stdout, out = _both(args, tmp_path)
    assert rc == 0
    if seed == 0:
        _both(["simu", "sv.txt", "cnv.txt", "calls.txt"], tmp_path)       # default chromosome is the empty string


def test_usage_and_errors(tmp_path):
    for args in ([], ["simu"], ["crest"], ["seeksv"], ["nonsense"], ["nonsense", "a", "b"], ["seeksv", "a"], ["simu", "a", "b", "c"]):
        got = []
        for exe in (REF, OURS):
            r = subprocess.run([exe] + args, capture_output=True, text=True, cwd=str(tmp_path))
            got.append((r.returncode, r.stdout, r.stderr.replace(exe, "PROG")))
        assert got[0] == got[1], args
    # unreadable input: message + exit 1
    got = []
    for exe in (REF, OURS):
        r = subprocess.run([exe, "seeksv", "missing.txt", "missing2.txt", str(tmp_path / "o")], capture_output=True, text=True, cwd=str(tmp_path))
        got.append((r.returncode, r.stdout, r.stderr))
    assert got[0] == got[1] and got[0][0] == 1
