#!/usr/bin/env python
"""Digests of the REAL reference's outputs on BASELINE.json's C2 workload (build container only: needs oracle/_ref/seeksv).

    python tests/golden/make_c2_digests.py      # ~2 minutes of single-core reference work

C2 = tools/svsim.cpp with the arguments bench.py uses (chr21-sized chromosome, 30x, 150 bp pairs, 500 planted events, seed
20261017): 9.2 M records, 2.7 GB uncompressed. The reference's outputs are 65 MB, so only their MD5s and sizes are committed
(tests/golden/c2/digests.json); the GPU test regenerates the same BAM with svsim on the box and compares digests. The realign
step is tools/minialign.cpp in both arms (deterministic; bwa is not on the GPU box).
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SEEKSV = os.path.join(ROOT, "oracle", "_ref", "seeksv")
BIN = os.path.join(ROOT, "seeksv_b200", "bin")
SVSIM_ARGS = ["--genome", "chr21:46709983", "--cov", "30", "--nsv", "500", "--seed", "20261017"]
# smaller stand-ins for the shapes of BASELINE.json's other configs (python tests/golden/make_c2_digests.py <workdir> c3mini ...):
#   c3mini: 24 contigs chr1..chr22, chrX, chrY (the name order chr1 < chr10 < chr11 ... < chr2 matters, quirk Q10), 30x
#   c5mini: human contig + HBV / HPV16 contigs at several thousand x (pileup cap, quirk Q12), virus-integration junctions
CONFIGS = {
    "c2": SVSIM_ARGS,
    "c3mini": ["--genome", ",".join("chr%s:%d" % (n, 300000 + 20000 * i) for i, n in enumerate(list(range(1, 23)) + ["X", "Y"])),
               "--cov", "30", "--nsv", "240", "--seed", "20261018"],
    "c5mini": ["--genome", "chr21:6000000", "--virus", "--cov", "30", "--nsv", "80", "--seed", "20261019"],
    # the shapes of configs 5 and 3 at the C2 size (round 2):
    #   c5: the C2 chromosome + HBV / HPV16 contigs at 6000x with 40 planted integrations (9.6 M records)
    #   c3w8: 24 contigs of 1.95 Mb (same total as C2): every one of 8 coordinate-range shards crosses chromosome boundaries
    "c5": ["--genome", "chr21:46709983", "--virus", "--cov", "30", "--nsv", "500", "--seed", "20261020"],
    "c3w8": ["--genome", ",".join("chr%s:1946249" % n for n in list(range(1, 23)) + ["X", "Y"]), "--cov", "30", "--nsv", "500",
             "--seed", "20261021"],
}
# config 4 at the C2 size: a 60x tumour and a 30x normal of the same donor (python tests/golden/make_c2_digests.py <workdir> c4)
PAIRS = {
    "c4": (["--genome", "chr21:46709983", "--cov", "60", "--nsv", "500", "--seed", "20261022", "--sample", "tumor"],
           ["--genome", "chr21:46709983", "--cov", "30", "--nsv", "500", "--seed", "20261022", "--sample", "normal"]),
}


def digest(data):
    return {"md5": hashlib.md5(data).hexdigest(), "bytes": len(data)}


def main():
    assert os.path.exists(SEEKSV), "oracle/build_ref.sh first"
    work = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="c2_")
    for name in (sys.argv[2:] or ["c2"]):
        if name in PAIRS:
            pair(work, name, *PAIRS[name])
        else:
            one(work, name, CONFIGS[name])


def pair(work, name, tumor_args, normal_args):
    """tumour: getclip -> minialign -> getsv; normal: getclip; somatic(normal.bam, normal.clip.gz, tumour.sv)"""
    out = {"tumor_args": tumor_args, "normal_args": normal_args, "reference": "oracle/_ref/seeksv (v1.2.3 built from the unmodified sources)"}
    pre = {}
    for kind, args in (("tumor", tumor_args), ("normal", normal_args)):
        pre[kind] = os.path.join(work, name + "." + kind)
        if not os.path.exists(pre[kind] + ".bam"):
            subprocess.run([os.path.join(BIN, "svsim"), "--out", pre[kind]] + args, check=True)
        ref = pre[kind] + ".ref"
        subprocess.run([SEEKSV, "getclip", "-o", ref, pre[kind] + ".bam"], check=True, stderr=subprocess.DEVNULL)
        for ext in (".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz"):
            with gzip.open(ref + ext, "rb") as f:
                out[kind + ext] = digest(f.read())
    ref = pre["tumor"] + ".ref"
    with open(ref + ".clip.sam", "wb") as o:
        subprocess.run([os.path.join(BIN, "minialign"), pre["tumor"] + ".fa", ref + ".clip.fq.gz"], check=True, stdout=o)
    out["tumor.clip.sam"] = digest(open(ref + ".clip.sam", "rb").read())
    r = subprocess.run([SEEKSV, "getsv", ref + ".clip.sam", pre["tumor"] + ".bam", ref + ".clip.gz", ref + ".sv", ref + ".unm"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    out["tumor getsv"] = {"sv": digest(open(ref + ".sv", "rb").read()), "stdout": digest(r.stdout)}
    subprocess.run([SEEKSV, "somatic", pre["normal"] + ".bam", pre["normal"] + ".ref.clip.gz", ref + ".sv", ref + ".somatic"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out["somatic"] = digest(open(ref + ".somatic", "rb").read())
    out["somatic rows"] = open(ref + ".somatic", "rb").read().count(b"\n")
    with open(os.path.join(HERE, "c2", name + ".digests.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print(json.dumps(out, indent=1))


def one(work, name, svsim_args):
    pre = os.path.join(work, name)
    if not os.path.exists(pre + ".bam"):
        subprocess.run([os.path.join(BIN, "svsim"), "--out", pre] + svsim_args, check=True)
    ref = os.path.join(work, name + ".ref")
    subprocess.run([SEEKSV, "getclip", "-o", ref, pre + ".bam"], check=True, stderr=subprocess.DEVNULL)
    out = {"svsim_args": svsim_args, "reference": "oracle/_ref/seeksv (v1.2.3 built from the unmodified sources)"}
    for ext in (".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz"):
        with gzip.open(ref + ext, "rb") as f:
            out[ext] = digest(f.read())
    with open(ref + ".clip.sam", "wb") as o:
        subprocess.run([os.path.join(BIN, "minialign"), pre + ".fa", ref + ".clip.fq.gz"], check=True, stdout=o)
    out["clip.sam"] = digest(open(ref + ".clip.sam", "rb").read())
    for tag, extra in (("getsv", []), ("getsv -n 0 -D", ["-n", "0", "-D"])):
        r = subprocess.run([SEEKSV, "getsv", *extra, ref + ".clip.sam", pre + ".bam", ref + ".clip.gz", ref + ".sv", ref + ".unm"],
                           check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        out[tag] = {"sv": digest(open(ref + ".sv", "rb").read()), "stdout": digest(r.stdout)}
    # somatic of the sample against itself
    subprocess.run([SEEKSV, "getsv", ref + ".clip.sam", pre + ".bam", ref + ".clip.gz", ref + ".sv", ref + ".unm"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([SEEKSV, "somatic", pre + ".bam", ref + ".clip.gz", ref + ".sv", ref + ".somatic"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out["somatic (self)"] = digest(open(ref + ".somatic", "rb").read())
    with open(os.path.join(HERE, "c2", "digests.json" if name == "c2" else name + ".digests.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
