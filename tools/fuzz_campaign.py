"""Seeded fuzz campaign on the CPU (build container only: needs /root/reference for bwa and oracle/_ref for the reference).

    python tools/fuzz_campaign.py [--seeds 100:140] [--records 1800] [--options 3] [--gpu]

For every seed: tests/fuzzgen.py writes an awkward BAM, the REAL reference (oracle/_ref/seeksv) runs getclip -> bwa mem ->
getsv (plain, `-n 0 -D`, `-B`) -> somatic(self), and then
  * the Python oracle (oracle/getclip_oracle.py, getsv_oracle.py) must reproduce every output byte for byte - this pins the
    checker on inputs beyond the committed fixtures;
  * the host layer of the product (`seeksv getsv -n 0 -D`, no BAM pass -> runs without a GPU) must reproduce the `-n 0 -D`
    outputs;
  * the rules of the device join (tools/clipjoin_sim.cpp) must build the host mirror's junction map from the same two files;
  * with --gpu (on a B200 box with oracle/_ref present; bwa's SAM is then replaced by tools/minialign) the product CLI runs the
    whole pipeline too.
Prints one line per seed and a summary; exit status 1 if anything differed. Test infrastructure, not product code.
"""
import argparse
import gzip
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import fuzzgen  # noqa: E402
from oracle import bamio, getclip_oracle, getsv_oracle  # noqa: E402

REF = os.environ.get("SEEKSV_REFERENCE", "/root/reference")
SEEKSV = os.path.join(ROOT, "oracle", "_ref", "seeksv")
BAMTOOL = os.path.join(ROOT, "oracle", "_ref", "bamtool")
CLI = os.path.join(ROOT, "seeksv_b200", "bin", "seeksv")
MINI = os.path.join(ROOT, "seeksv_b200", "bin", "minialign")
JOIN_SIM = os.path.join(ROOT, "seeksv_b200", "bin", "clipjoin_sim")


def text(p):
    with open(p, "rb") as f:
        return f.read().decode("latin-1")


def zcat(p):
    with gzip.open(p, "rb") as f:
        return f.read().decode("latin-1")


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw)


def one_seed(seed, records, work, bwa, gpu, n_opts=0, edge=False, connect_n=0):
    bad = []
    bam = os.path.join(work, "f.sort.bam")
    _, _, genome = fuzzgen.write(bam, seed, records, edge)
    subprocess.run([BAMTOOL, "index", bam], check=True)
    fa = os.path.join(work, "f.fa")
    fuzzgen.write_fasta(genome, fa)
    pre = os.path.join(work, "ref")
    r = run([SEEKSV, "getclip", "-o", pre, bam])
    if r.returncode != 0:
        return ["reference getclip failed (%d)" % r.returncode]
    ref = {e: zcat(pre + e) for e in (".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz")}
    sam = os.path.join(work, "clip.sam")
    if bwa:
        subprocess.run([bwa, "index", fa], check=True, capture_output=True)
        with open(sam, "w") as o:
            subprocess.run([bwa, "mem", fa, pre + ".clip.fq.gz"], check=True, stdout=o, stderr=subprocess.DEVNULL)
    else:
        with open(sam, "w") as o:
            subprocess.run([MINI, fa, pre + ".clip.fq.gz"], check=True, stdout=o)
    variants = {"": [], ".n0D": ["-n", "0", "-D"]}
    for tag, extra in list(variants.items()):
        r = run([SEEKSV, "getsv", *extra, sam, bam, pre + ".clip.gz", pre + tag + ".sv", pre + ".unm"])
        if r.returncode != 0:
            return ["reference getsv%s failed (%d)" % (tag, r.returncode)]
        ref["sv" + tag], ref["out" + tag] = text(pre + tag + ".sv"), r.stdout
    r = run([SEEKSV, "getsv", "-B", pre + ".sv", sam, bam, pre + ".clip.gz", pre + ".B.sv", pre + ".unm"])
    ref["sv.B"], ref["out.B"] = (text(pre + ".B.sv"), r.stdout) if r.returncode == 0 else (None, None)
    r = run([SEEKSV, "somatic", bam, pre + ".clip.gz", pre + ".sv", pre + ".somatic"])
    ref["somatic"] = text(pre + ".somatic") if r.returncode == 0 else None

    # --- the oracle against the reference
    h, recs = bamio.read_bam(bam)
    got = getclip_oracle.getclip(h, recs)
    for g, e in zip(got, (".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz")):
        if g != ref[e]:
            bad.append("oracle getclip " + e)
    ch, ca = bamio.read_alignments(sam)
    sv, out = getsv_oracle.getsv(h, recs, ref[".clip.gz"], ch, ca)
    if sv != ref["sv"]:
        bad.append("oracle getsv .sv")
    if out != ref["out"]:
        bad.append("oracle getsv stdout")
    if ref["somatic"] is not None:
        if getsv_oracle.somatic(h, recs, ref[".clip.gz"], ref["sv"]) != ref["somatic"]:
            bad.append("oracle somatic")

    # --- the product's host layer (no BAM pass, no GPU) against the reference
    if os.path.exists(CLI):
        r = run([CLI, "getsv", "-n", "0", "-D", sam, bam, pre + ".clip.gz", os.path.join(work, "host.sv"), os.path.join(work, "host.unm")])
        if r.returncode != 0:
            bad.append("host getsv -n0 -D exit %d" % r.returncode)
        else:
            if text(os.path.join(work, "host.sv")) != ref["sv.n0D"]:
                bad.append("host -n0 -D .sv")
            if r.stdout != ref["out.n0D"]:
                bad.append("host -n0 -D stdout")

    # --- the rules of the device join (csrc/clipjoin_core.h, run serially by tools/clipjoin_sim.cpp) against the host mirror, on the
    #     reference's own clip.gz and the aligner's SAM: also with every chunk's entry guess wrong (repair rounds)
    if os.path.exists(JOIN_SIM):
        for extra in ([], ["--wrong-guesses"]):
            r = run([JOIN_SIM, sam, pre + ".clip.gz"] + extra)
            if r.returncode != 0 or not r.stdout.startswith(("OK", "SKIP")):
                bad.append("clip_join rules %s: %s" % (" ".join(extra), r.stdout.strip()[:80]))

    # --- the whole product pipeline (B200 only)
    if gpu:
        p = os.path.join(work, "b200")
        r = run([CLI, "getclip", "-o", p, bam])
        if r.returncode != 0:
            bad.append("b200 getclip exit %d: %s" % (r.returncode, r.stderr[-200:]))
        else:
            for e in (".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz"):
                if zcat(p + e) != ref[e]:
                    bad.append("b200 getclip " + e)
            r = run([CLI, "getsv", sam, bam, p + ".clip.gz", p + ".sv", p + ".unm"])
            if r.returncode != 0 or text(p + ".sv") != ref["sv"] or r.stdout != ref["out"]:
                bad.append("b200 getsv")
            if ref["sv.B"] is not None:
                r = run([CLI, "getsv", "-B", pre + ".sv", sam, bam, p + ".clip.gz", p + ".B.sv", p + ".unm"])
                if r.returncode != 0 or text(p + ".B.sv") != ref["sv.B"] or r.stdout != ref["out.B"]:
                    bad.append("b200 getsv -B")
            if ref["somatic"] is not None:
                r = run([CLI, "somatic", bam, pre + ".clip.gz", pre + ".sv", p + ".somatic"])
                if r.returncode != 0 or text(p + ".somatic") != ref["somatic"]:
                    bad.append("b200 somatic")
    # --- the optional junction sources: -F <connected read-through reads> (FindJunction, process_bwasw.cpp:5-227), -w, -B
    if connect_n:
        import make_golden
        connect = os.path.join(work, "connect.sam")
        make_golden.build_connect_sam(bam, connect, seed=seed, n_reads=connect_n)
        cset = bamio.read_alignments(connect)
        w = random.Random(seed).choice([1, 1, 30, 60])
        for tag, argv, kw in (("-F", ["-F", connect, "-w", str(w), "-n", "0", "-D"], dict(connect=cset, connect_min_mapq=w, pairs_used=0, output_depth=False)),
                              ("-F full", ["-F", connect], dict(connect=cset)),
                              ("-F -B", ["-F", connect, "-B", pre + ".sv", "-n", "0", "-D"],
                               dict(connect=cset, seed_text=ref["sv"], pairs_used=0, output_depth=False))):
            out_sv = os.path.join(work, "cF.sv")
            r = run([SEEKSV, "getsv", *argv, sam, bam, pre + ".clip.gz", out_sv, pre + ".unm"])
            if r.returncode != 0:
                continue
            want_sv, want_out = text(out_sv), r.stdout
            sv, out = getsv_oracle.getsv(h, recs, ref[".clip.gz"], ch, ca, **kw)
            if sv != want_sv or out != want_out:
                bad.append("oracle getsv " + tag)
            if os.path.exists(CLI) and (gpu or "-D" in argv):
                r2 = run([CLI, "getsv", *argv, sam, bam, pre + ".clip.gz", os.path.join(work, "cF2.sv"), os.path.join(work, "cF2.unm")])
                if r2.returncode != 0 or text(os.path.join(work, "cF2.sv")) != want_sv or r2.stdout != want_out:
                    bad.append(("b200 getsv " if "-D" not in argv else "host getsv ") + tag)

    # --- random option vectors (Appendix E of SURVEY.md): reference vs oracle (full getsv, getclip) and vs the host layer
    rng = random.Random(seed * 7919 + 1)
    for k in range(n_opts):
        # getclip -t / -q / -s
        t, q, sl = rng.choice([0.5, 0.8, 0.85, 0.9, 0.95, 1.0]), rng.choice([0, 1, 10, 20, 30, 60]), rng.random() < 0.5
        po = os.path.join(work, "opt%d" % k)
        r = run([SEEKSV, "getclip", "-t", str(t), "-q", str(q)] + (["-s"] if sl else []) + ["-o", po, bam])
        if r.returncode == 0:
            want = [zcat(po + e) for e in (".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz")]
            if list(getclip_oracle.getclip(h, recs, limit=t, min_mapq=q, save_low_quality=sl)) != want:
                bad.append("oracle getclip -t %s -q %d%s" % (t, q, " -s" if sl else ""))
        # getsv: every numeric option that has an effect
        o = dict(l=rng.choice([0, 5, 20, 50, 90]), q=rng.choice([0, 1, 20, 30]), b=rng.choice([1, 2, 3, 5]), d=rng.choice([0, 10, 50, 500]),
                 T=rng.choice([0, 5, 50]), m=rng.choice([0, 30, 60]), i=rng.choice([0, 1, 3]), L=rng.choice([1, 20, 200, 1000]),
                 e=rng.choice([0, 0, 1, 2]), f=rng.choice([0, 0.05, 0.1, 0.5]))
        argv = []
        for name, v in o.items():
            argv += ["-" + name, str(v)]
        r = run([SEEKSV, "getsv", *argv, sam, bam, pre + ".clip.gz", po + ".sv", po + ".unm"])
        if r.returncode == 0:
            sv, out = getsv_oracle.getsv(h, recs, ref[".clip.gz"], ch, ca, flank=o["l"], min_mapq=o["q"], min_clip_sum=o["b"],
                                         min_dist=o["d"], max_micro=o["T"], min_seq_len=o["m"], max_indel=o["i"], flank_len=o["L"],
                                         min_pairs=o["e"], freq=o["f"])
            if sv != text(po + ".sv") or out != r.stdout:
                bad.append("oracle getsv " + " ".join(argv))
        # the host layer with the same options, BAM passes off
        hargv = [a for a in argv] + ["-n", "0", "-D"]
        r = run([SEEKSV, "getsv", *hargv, sam, bam, pre + ".clip.gz", po + ".h.sv", po + ".unm"])
        if r.returncode == 0 and os.path.exists(CLI):
            r2 = run([CLI, "getsv", *hargv, sam, bam, pre + ".clip.gz", po + ".h2.sv", po + ".unm2"])
            if r2.returncode != 0 or text(po + ".h2.sv") != text(po + ".h.sv") or r2.stdout != r.stdout:
                bad.append("host getsv " + " ".join(hargv))
        # somatic -t / -q / -l / -m (the sample against itself)
        so = dict(t=rng.choice([0.5, 0.8, 0.9, 1.0]), q=rng.choice([0, 20, 30]), l=rng.choice([0, 10, 30, 89]), m=rng.choice([1, 10, 40]),
                  n=rng.choice([5000000, 5000000, 0, 99999]))
        sargv = []
        for name, v in so.items():
            sargv += ["-" + name, str(v)]
        r = run([SEEKSV, "somatic", *sargv, bam, pre + ".clip.gz", pre + ".sv", po + ".somatic"])
        if r.returncode == 0:
            if getsv_oracle.somatic(h, recs, ref[".clip.gz"], ref["sv"], rate=so["t"], min_mapq=so["q"], offset=so["l"],
                                    min_len=so["m"], pairs_used=so["n"]) != text(po + ".somatic"):
                bad.append("oracle somatic " + " ".join(sargv))
            if gpu and os.path.exists(CLI):
                r2 = run([CLI, "somatic", *sargv, bam, pre + ".clip.gz", pre + ".sv", po + ".g.somatic"])
                if r2.returncode != 0 or text(po + ".g.somatic") != text(po + ".somatic"):
                    bad.append("b200 somatic " + " ".join(sargv))
        if gpu and os.path.exists(CLI):
            r = run([SEEKSV, "getsv", *argv, sam, bam, pre + ".clip.gz", po + ".sv", po + ".unm"])
            r2 = run([CLI, "getsv", *argv, sam, bam, pre + ".clip.gz", po + ".g.sv", po + ".unm2"])
            if r.returncode == 0 and (r2.returncode != 0 or text(po + ".g.sv") != text(po + ".sv") or r2.stdout != r.stdout):
                bad.append("b200 getsv " + " ".join(argv))
    n_sv = ref["sv"].count("\n")
    return bad, n_sv, len(recs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="100:120")
    ap.add_argument("--records", type=int, default=1800)
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--options", type=int, default=0, help="random option vectors per seed (getclip -t/-q/-s, getsv -l..-f)")
    ap.add_argument("--connect", type=int, default=0, help="also run getsv -F / -w / -B with this many connected read-through reads")
    ap.add_argument("--contigs", default="", help="name:length,... instead of fuzzgen's chrB / chrA / virus (map orders by NAME matter)")
    ap.add_argument("--edge", action="store_true", help="fuzzgen edge mode: breakpoints at the contig ends, clipped parts of 320 bases")
    ap.add_argument("--keep", action="store_true", help="keep the work directory of failing seeds")
    a = ap.parse_args()
    lo, hi = (int(x) for x in a.seeds.split(":"))
    if a.contigs:
        fuzzgen.CONTIGS = [(c.split(":")[0], int(c.split(":")[1])) for c in a.contigs.split(",")]
    assert os.path.exists(SEEKSV) and os.path.exists(BAMTOOL), "oracle/build_ref.sh first"
    top = tempfile.mkdtemp(prefix="fuzzcamp_")
    bwa = None
    if os.path.exists(os.path.join(REF, "example", "bin", "bwa")):
        bwa = os.path.join(top, "bwa")
        shutil.copy(os.path.join(REF, "example", "bin", "bwa"), bwa)
        os.chmod(bwa, 0o755)
    failed = 0
    for seed in range(lo, hi):
        work = os.path.join(top, "s%d" % seed)
        os.makedirs(work)
        res = one_seed(seed, a.records, work, bwa, a.gpu, a.options, a.edge, a.connect)
        if isinstance(res, list):
            print("seed %d: SKIP %s" % (seed, res[0]), flush=True)
            shutil.rmtree(work)
            continue
        bad, n_sv, n_rec = res
        print("seed %d: %d records, %d sv lines: %s" % (seed, n_rec, n_sv, "ok" if not bad else "DIFF " + "; ".join(bad)), flush=True)
        if bad:
            failed += 1
            if a.keep:
                continue
        shutil.rmtree(work)
    if not (failed and a.keep):
        shutil.rmtree(top, ignore_errors=True)
    else:
        print("kept:", top)
    print("%d of %d seeds differed" % (failed, hi - lo))
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
