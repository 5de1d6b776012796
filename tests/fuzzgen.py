"""Seeded generator of awkward coordinate-sorted BAMs for the parity tests (test infrastructure).

What it stresses, beyond the simulators: every CIGAR op of "MIDNSHP=X" in random order, hard clips on either end, reads of
8-40 kb (records longer than the 16 KiB chunks the CUDA walkers cut the stream into), soft clips on one or both ends with
and without the XC tag, aux fields of every BAM type in front of XC (floats and doubles make the linked libbam lose its
place in the aux block - the oracle and the kernel follow it), all eleven flag bits, mapQ 0-60, unmapped-branch
records with repeated names, sequences with IUPAC codes and lower case, records without sequence, and clusters of reads
that share a breakpoint with 0-25 % disagreement (so that the 0.9 merge threshold cuts both ways).

Avoided on purpose (undefined behaviour in the reference): soft-clipped reads without a quality
string, mapped records with an empty CIGAR.
"""
import os
import random
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bamio  # noqa: E402

CONTIGS = [("chrB", 60000), ("chrA", 30000), ("virus", 4000)]   # names out of lexicographic order on purpose


def _aux(rng, xc=None):
    out = b""
    for _ in range(rng.randrange(0, 4)):
        tag = bytes(rng.choice(b"ABCDEFGHIJKLMNOPQRSTUVWYZ") for _ in range(2))   # never 'X?' so XC stays unique
        t = rng.choice("cCsSiIAZfdHB")
        if t == "c":
            out += tag + b"c" + struct.pack("<b", rng.randrange(-128, 128))
        elif t == "C":
            out += tag + b"C" + struct.pack("<B", rng.randrange(256))
        elif t == "s":
            out += tag + b"s" + struct.pack("<h", rng.randrange(-30000, 30000))
        elif t == "S":
            out += tag + b"S" + struct.pack("<H", rng.randrange(65536))
        elif t == "i":
            out += tag + b"i" + struct.pack("<i", rng.randrange(-10 ** 6, 10 ** 6))
        elif t == "I":
            out += tag + b"I" + struct.pack("<I", rng.randrange(2 ** 32))
        elif t == "A":
            out += tag + b"A" + bytes([rng.randrange(33, 127)])
        elif t == "Z":
            out += tag + b"Z" + bytes(rng.randrange(33, 127) for _ in range(rng.randrange(0, 30))) + b"\0"
        elif t == "f":   # (the linked libbam does not know how to step over floats and doubles: see oracle/bamio.py:aux_walk)
            out += tag + b"f" + struct.pack("<f", rng.random())
        elif t == "d":
            out += tag + b"d" + struct.pack("<d", rng.random())
        elif t == "H":
            out += tag + b"H" + bytes(rng.choice(b"0123456789ABCDEF") for _ in range(2 * rng.randrange(0, 8))) + b"\0"
        else:
            sub = rng.choice("cCsSiIf")
            n = rng.randrange(0, 6)
            out += tag + b"B" + sub.encode() + struct.pack("<i", n) + bytes(rng.randrange(256) for _ in range(n * {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]))
    if xc is not None:
        kind = rng.choice("cCsSiI" if xc < 128 else "sSiI" if xc < 32768 else "SiI")
        out += b"XC" + kind.encode() + struct.pack({"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[kind], xc)
    return out


def generate(seed, n_records=2500, edge=False):
    """edge=True (tools/fuzz_campaign.py --edge; the committed fixtures use edge=False and its random stream is untouched):
    breakpoints and the places the clipped parts come from crowd at the first and last 250 bases of every contig (window clamps,
    unsigned flank ranges, the first keys of the range maps), and one clipped part in ten is 320 bases long (read names of
    255+ characters in the realigner's SAM)."""
    rng = random.Random(seed)
    genome = ["".join(rng.choice("ACGT") for _ in range(ln)) for _, ln in CONTIGS]
    foreign = {}   # breakpoint -> clipped-away sequence shared by the reads of that breakpoint

    def far_segment():   # 60 bases of the genome elsewhere (either strand), so that the realigner can place the clipped parts
        t = rng.randrange(len(CONTIGS))
        n = 60
        if edge:
            n = 320 if rng.random() < 0.1 else 60
            q = rng.choice([rng.randrange(0, 200), rng.randrange(CONTIGS[t][1] - n - 200, CONTIGS[t][1] - n), rng.randrange(0, CONTIGS[t][1] - n)])
        else:
            q = rng.randrange(0, CONTIGS[t][1] - 60)
        seg = genome[t][q:q + n]
        if rng.random() < 0.5:
            seg = seg[::-1].translate(str.maketrans("ACGT", "TGCA"))
        return seg
    recs = []
    for tid, (_, ln) in enumerate(CONTIGS):
        n_here = n_records * ln // sum(l for _, l in CONTIGS)
        if edge:
            breakpoints = sorted(rng.choice([rng.randrange(1, 250), rng.randrange(ln - 250, ln - 1), rng.randrange(200, ln - 200)])
                                 for _ in range(max(3, n_here // 40)))
        else:
            breakpoints = sorted(rng.randrange(200, ln - 200) for _ in range(max(3, n_here // 40)))
        for i in range(n_here):
            long_read = rng.random() < 0.01
            aligned = rng.randrange(8000, 40000) if long_read else rng.randrange(30, 151)
            use_bp = rng.random() < 0.45
            if use_bp:
                bp = rng.choice(breakpoints)
                side = rng.choice("LR")
                pos0 = bp if side == "L" else max(0, bp - aligned)
                aligned = min(aligned, ln - pos0) if side == "L" else bp - pos0
            else:
                pos0 = rng.randrange(0, ln - 1)
                aligned = min(aligned, ln - pos0)
                side = rng.choice("LRB-") if rng.random() < 0.3 else "-"
            if aligned < 5:
                continue
            # aligned part: random op mix that consumes `aligned` reference bases
            # (first and last op are M: libbam's pileup asserts on alignments that begin or end with other ops)
            ops, ref_left, read_seq = [], aligned, []
            rp = pos0
            tail_m = rng.randrange(1, 4)
            ref_left -= tail_m
            while ref_left > 0:
                op = rng.choices("MIDN=XP", weights=[60, 6, 6, 2, 8, 6, 1])[0]
                if not ops:
                    op = "M"
                k = min(ref_left, rng.randrange(1, 60) if not long_read else rng.randrange(50, 3000))
                if op in "M=X":
                    read_seq.append(genome[tid][rp:rp + k])
                    rp += k
                    ref_left -= k
                elif op in "DN":
                    k = min(k, ref_left - 1) if ref_left > 1 else 0
                    if k == 0:
                        continue
                    rp += k
                    ref_left -= k
                elif op == "I":
                    k = rng.randrange(1, 8)
                    read_seq.append("".join(rng.choice("ACGT") for _ in range(k)))
                else:
                    k = rng.randrange(1, 4)
                ops.append((k, op))
            ops.append((tail_m, "M"))
            read_seq.append(genome[tid][rp:rp + tail_m])
            body = "".join(read_seq)
            # disagreements with the genome (so that clusters sometimes refuse a read)
            err = rng.choice([0.0, 0.0, 0.02, 0.1, 0.25])
            body = "".join(rng.choice("ACGTN") if rng.random() < err else c for c in body)
            lclip = rclip = ""
            hl = hr = 0
            if side in "LB":
                key = (tid, pos0, "L")
                if key not in foreign:
                    foreign[key] = far_segment()
                f = foreign[key]
                k = rng.randrange(1, len(f) + 1)
                lclip = f[len(f) - k:]
            if side in "RB":
                key = (tid, pos0 + aligned, "R")
                if key not in foreign:
                    foreign[key] = far_segment()
                f = foreign[key]
                k = rng.randrange(1, len(f) + 1)
                rclip = f[:k]
            if rng.random() < 0.04:
                hl = rng.randrange(1, 50)
            if rng.random() < 0.04:
                hr = rng.randrange(1, 50)
            cig = ""
            if hl:
                cig += "%dH" % hl
            if lclip:
                cig += "%dS" % len(lclip)
            merged = []
            for k, op in ops:   # merge neighbours of the same op (libbam keeps them apart, but keep the file tidy)
                if merged and merged[-1][1] == op:
                    merged[-1] = (merged[-1][0] + k, op)
                else:
                    merged.append((k, op))
            cig += "".join("%d%s" % (k, op) for k, op in merged)
            if rclip:
                cig += "%dS" % len(rclip)
            if hr:
                cig += "%dH" % hr
            seq = lclip + body + rclip
            if rng.random() < 0.05:
                seq = "".join(rng.choice("RYKMSWBDHVN") if rng.random() < 0.05 else c for c in seq)
            if rng.random() < 0.05:
                seq = seq.lower()
            clipped = bool(lclip or rclip)
            if not clipped and rng.random() < 0.02:
                qual = "*"
            else:
                qual = "".join(chr(33 + min(60, max(0, int(rng.gauss(35, 12))))) for _ in seq)
            flag = 0
            for bit, p in ((1, .9), (2, .7), (16, .5), (32, .5), (64, .5), (128, .5), (256, .04), (512, .03), (1024, .04), (2048, .03)):
                if rng.random() < p:
                    flag |= bit
            unmapped_branch = rng.random() < 0.06
            if unmapped_branch:
                flag |= rng.choice([4, 8, 12])
                name = "um%d" % rng.randrange(0, 40)          # few names: repeats, same-end duplicates, triples
            else:
                name = "r%d_%d" % (tid, i)
            mapq = rng.choice([0, 0, 1, 5, 19, 20, 21, 29, 30, 37, 60, 60, 60])
            mtid = tid if rng.random() < 0.85 else rng.randrange(-1, len(CONTIGS))
            mpos = max(0, pos0 + rng.randrange(-900, 900)) if mtid == tid else rng.randrange(0, 3000)
            isize = rng.choice([0, mpos - pos0, rng.randrange(-1000, 1000), rng.randrange(200, 800), 50000, -50000])
            xc = None
            if lclip and rclip and rng.random() < 0.5:
                xc = rng.choice([0, len(seq) - len(rclip), len(lclip), rng.randrange(1, 200)])
            aux = _aux(rng, xc)
            while clipped and bamio.aux_walk(aux, b"XC")[1]:   # a walk that leaves the record is undefined in the reference
                aux = _aux(rng, xc)
            recs.append(bamio.make_rec(name, flag, tid, pos0, mapq, cig, mtid, mpos, isize, seq, qual, aux))
    # a handful of records without sequence, and unmapped reads at the end of the file
    for _ in range(5):
        tid = rng.randrange(len(CONTIGS))
        recs.append(bamio.make_rec("noseq%d" % _, 0, tid, rng.randrange(0, 1000), 30, "10M", -1, -1, 0, "", "", b""))
    order = {id(r): k for k, r in enumerate(recs)}
    recs.sort(key=lambda r: (r.tid, r.pos, order[id(r)]))
    for k in range(6):
        recs.append(bamio.make_rec("tail%d" % (k // 2), 4 | 1 | (64 if k & 1 else 128), -1, -1, 0, "*", -1, -1, 0,
                                   "ACGTNACGT"[:5 + k % 3], "IIIIIIIII"[:5 + k % 3], b""))
    header = bamio.Header([c for c, _ in CONTIGS], [l for _, l in CONTIGS], "@HD\tVN:1.0\tSO:coordinate\n")
    return header, recs, genome


def write(path, seed, n_records=2500, edge=False):
    h, recs, genome = generate(seed, n_records, edge)
    bamio.write_bam(path, h, recs)
    return h, recs, genome


def write_fasta(genome, path):
    with open(path, "w") as f:
        for (name, _), seq in zip(CONTIGS, genome):
            f.write(">%s\n" % name)
            for i in range(0, len(seq), 60):
                f.write(seq[i:i + 60] + "\n")


if __name__ == "__main__":
    h, recs, _ = write(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    print(len(recs), "records")
