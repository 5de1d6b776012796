"""SAM text input (getclip accepts it for any file name not ending in .bam, getsv for the realigned clips): the host-side
conversion (svb_sam_to_stream == what svb_bam_open / getsv use) against the linked libbam's own parser (`bamtool sam2bam` =
samopen(fn, "r") + samwrite, the code path behind clip_reads.h:375 / getsv.h:445). Byte-for-byte on the uncompressed stream:
header, bin, aux integer narrowing, '*' fields, '=' mate reference, IUPAC and lower-case bases, names of 255+ characters.
Left out: `d` (double) aux values and float arrays (`B:f`) - this libbam stores garbage for them (0 for 1.5, 5.0f for -3.25:
probed) -, empty Z/H
values and lines with CIGAR ops `=`/`X` (it aborts on both). CPU only."""
import gzip
import os
import random
import subprocess

import pytest

from conftest import GOLDEN, ROOT
from oracle import bamio

BAMTOOL = os.path.join(ROOT, "oracle", "_ref", "bamtool")
pytestmark = pytest.mark.skipif(not os.path.exists(BAMTOOL), reason="no oracle/_ref (oracle/build_ref.sh)")


def _libbam_stream(sam, tmp_path):
    bam = str(tmp_path / "x.bam")
    subprocess.run([BAMTOOL, "sam2bam", sam, bam], check=True, capture_output=True)
    with gzip.open(bam, "rb") as f:
        return f.read()


@pytest.mark.parametrize("rel", ["example/cancer.clip.sam", "micro/tumor.clip.sam", "micro/tumor.connect.sam", "fuzz/f11.clip.sam",
                                 "fuzz/f106.clip.sam"])
def test_golden_sam_files_convert_like_libbam(rel, tmp_path):
    from seeksv_b200 import lib
    sam = os.path.join(GOLDEN, rel)
    got, first = lib.sam_to_stream(sam)
    want = _libbam_stream(sam, tmp_path)
    assert got == want
    assert got[:4] == b"BAM\1" and 0 < first < len(got)
    assert bamio.sam_to_stream(sam) == want          # the oracle's reader is pinned the same way


def _random_sam(rng, n):
    contigs = [("chrB", 50000), ("chrA", 30000), ("virus", 4000)]
    lines = ["@HD\tVN:1.4\tSO:unsorted"] + ["@SQ\tSN:%s\tLN:%d" % c for c in contigs] + ["@PG\tID:x\tPN:x"]
    for i in range(n):
        unmapped = rng.random() < 0.1
        tid = rng.randrange(len(contigs))
        l = rng.choice([0, 1, 7, 36, 100, 151])
        ops = []
        if not unmapped and l:
            left = l
            if rng.random() < 0.4 and left > 2:
                k = rng.randrange(1, left // 2 + 1)
                ops.append("%d%s" % (k, rng.choice("SH")))
                left -= k if ops[-1][-1] == "S" else 0
            m = max(1, left - rng.randrange(0, max(1, left // 3)))
            ops.append("%dM" % m)
            left -= m
            if left > 0 and rng.random() < 0.5:
                ops += ["%dD" % rng.randrange(1, 20), "%dI" % 1] if left > 1 else []
                left -= 1 if left > 1 else 0
            if left > 0:
                ops.append("%d%s" % (left, rng.choice("SM")))
        cigar = "".join(ops) if ops else "*"
        qlen = sum(int(x[:-1]) for x in ops if x[-1] in "MIS") if ops else l
        seq = "".join(rng.choice("ACGTNacgtnRYKM") for _ in range(qlen)) or "*"
        qual = "*" if seq == "*" or rng.random() < 0.1 else "".join(chr(rng.randrange(33, 74)) for _ in range(len(seq)))
        flag = rng.choice([0, 16, 99, 147, 83, 163, 256, 272, 2048, 2064, 1024, 512]) | (4 if unmapped else 0)
        rname = "*" if unmapped and rng.random() < 0.5 else contigs[tid][0]
        pos = 0 if rname == "*" else rng.randrange(1, contigs[tid][1])
        rnext = rng.choice(["*", "=", contigs[rng.randrange(len(contigs))][0]]) if rname != "*" else "*"
        pnext = 0 if rnext == "*" else rng.randrange(1, 4000)
        aux = []
        for _ in range(rng.randrange(0, 5)):
            tag = rng.choice("ABNXYZ") + rng.choice("CMST012")
            t = rng.choice("iiiAZfH")
            if t == "i":
                v = rng.choice([0, 1, -1, 127, 128, 255, 256, -128, -129, 32767, 32768, 65535, 65536, -32768, -32769, 2 ** 31 - 1,
                                -2 ** 31, 2 ** 32 - 1, rng.randrange(-10 ** 6, 10 ** 6)])
                aux.append("%s:i:%d" % (tag, v))
            elif t == "A":
                aux.append("%s:A:%s" % (tag, chr(rng.randrange(33, 127))))
            elif t == "Z":
                aux.append("%s:Z:%s" % (tag, "".join(chr(rng.randrange(33, 127)) for _ in range(rng.randrange(1, 20)))))   # (libbam aborts on an empty Z value)
            elif t == "f":
                aux.append("%s:f:%s" % (tag, rng.choice(["0.5", "1e-3", "-2.25", "3"])))
            else:
                aux.append("%s:H:%s" % (tag, "".join(rng.choice("0123456789ABCDEF") for _ in range(2 * rng.randrange(1, 6)))))
        name = rng.choice(["r%d" % i, "ACGTTGCAAC", "q" * rng.randrange(1, 60)])
        fields = [name, str(flag), rname, str(pos), str(rng.choice([0, 1, 20, 60, 255])), cigar, rnext, str(pnext),
                  str(rng.randrange(-1000, 1000)), seq, qual] + aux
        lines.append("\t".join(fields))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_sam_text_converts_like_libbam(seed, tmp_path):
    from seeksv_b200 import lib
    sam = str(tmp_path / "r.sam")
    with open(sam, "w") as f:
        f.write(_random_sam(random.Random(seed), 400))
    want = _libbam_stream(sam, tmp_path)
    got, first = lib.sam_to_stream(sam)
    if got != want:
        i = next(k for k in range(min(len(got), len(want))) if got[k] != want[k])
        raise AssertionError("streams differ at byte %d: libbam %r, ours %r" % (i, want[max(0, i - 24):i + 24], got[max(0, i - 24):i + 24]))
    assert bamio.sam_to_stream(sam) == want


EDGE_LINES = {
    "unknown reference name": "r1\t0\tchrZ\t100\t60\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII",
    "mapped flag without a reference": "r1\t0\t*\t100\t60\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII",
    "unknown mate reference": "r1\t0\tchrA\t100\t60\t10M\tchrQ\t5\t0\tACGTACGTAC\tIIIIIIIIII",
    "mapped flag without a CIGAR": "r1\t16\tchrA\t100\t60\t*\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII",
    "typed array": "r1\t0\tchrA\t100\t60\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII\tXB:B:i,1,-2,3\tXS:B:S,7,65535\tXC:i:5",
    "hexadecimal flag": "r1\t0x10\tchrA\t100\t60\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII",
    "flag letters": "r1\tpr1\tchrA\t100\t60\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII",
    "signed and oversized integers": "r1\t0\tchrA\t100\t60\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII\tXC:i:+5\tXD:i:4294967296\tXE:i:-128",
    "position 0 and mapQ 300": "r1\t0\tchrA\t0\t300\t10M\t*\t0\t0\tAC.TACGTAC\tIIIIIIIIII",
    "name of 254 characters": "q" * 254 + "\t0\tchrA\t10\t30\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII",
    "name of 255 characters": "q" * 255 + "\t0\tchrA\t10\t30\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII",
    "name of 300 characters": "q" * 300 + "\t0\tchrA\t10\t30\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII",
    "padding and skip ops": "r1\t0\tchrA\t10\t30\t3M2P2M200N5M\t=\t700\t-5\tACGTACGTAC\t*",
    "unmapped with a position": "r1\t4\tchrA\t10\t0\t*\t=\t10\t0\tACGTACGTAC\tIIIIIIIIII",
    "carriage return": "r1\t0\tchrA\t10\t30\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII\r",
}


@pytest.mark.parametrize("what", sorted(EDGE_LINES))
def test_edge_lines_convert_like_libbam(what, tmp_path):
    from seeksv_b200 import lib
    sam = str(tmp_path / "e.sam")
    with open(sam, "w", newline="") as f:
        f.write("@SQ\tSN:chrA\tLN:30000\n@SQ\tSN:chrB\tLN:50000\n" + EDGE_LINES[what] + "\n")
    want = _libbam_stream(sam, tmp_path)
    got, _ = lib.sam_to_stream(sam)
    assert got == want, what
    assert bamio.sam_to_stream(sam) == want, what
