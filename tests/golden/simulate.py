"""Seeded micro-scale read simulator for parity fixtures (test infrastructure).

Produces coordinate-sorted SAM text for a tumour-like sample over a small random genome with planted
structural variants, written as *already aligned* records (split reads -> soft clips, discordant
mates, duplicates, unmapped mates, hard-clipped supplementary records, a few =/X/N/I/D CIGAR ops
and XC tags) so that every branch of the reference's getclip/getsv/somatic path is exercised - the
bundled example only has three deletions.

Contig names are chosen so that name order != tid order.
"""
from __future__ import annotations

import random
from typing import List, Tuple

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def rc(s: str) -> str:
    return "".join(COMP[c] for c in reversed(s))


def reg2bin(beg: int, end: int) -> int:
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


class Genome:
    def __init__(self, rng: random.Random, contigs: List[Tuple[str, int]]):
        self.names = [c for c, _ in contigs]
        self.seq = {c: "".join(rng.choice("ACGT") for _ in range(n)) for c, n in contigs}


# a donor chromosome is a list of (contig, start0, end0, strand) reference segments
def donor_sequence(g: Genome, segs) -> str:
    out = []
    for c, s, e, st in segs:
        piece = g.seq[c][s:e]
        out.append(piece if st == "+" else rc(piece))
    return "".join(out)


def locate(segs, off: int):
    """segment index and offset inside it for a donor coordinate"""
    for i, (c, s, e, st) in enumerate(segs):
        n = e - s
        if off < n:
            return i, off
        off -= n
    raise IndexError


def align_read(g: Genome, segs, start: int, length: int, rng: random.Random):
    """Reference alignment of donor[start:start+length] read in donor-forward orientation.
    Returns list of pieces (contig, refpos0, reflen, strand, qbeg, qend) in query order."""
    pieces = []
    q = 0
    i, off = locate(segs, start)
    while q < length:
        c, s, e, st = segs[i]
        n = min(length - q, (e - s) - off)
        if st == "+":
            pieces.append((c, s + off, n, "+", q, q + n))
        else:
            pieces.append((c, e - off - n, n, "-", q, q + n))
        q += n
        i += 1
        off = 0
    return pieces


def make_records(g: Genome, segs, rng: random.Random, coverage: float, read_len: int, isize_mu: int,
                 isize_sd: int, tag: str, opts) -> List[dict]:
    donor = donor_sequence(g, segs)
    n_pairs = int(len(donor) * coverage / (2 * read_len))
    recs = []
    for k in range(n_pairs):
        isz = max(read_len + 10, int(rng.gauss(isize_mu, isize_sd)))
        if isz >= len(donor):
            continue
        s = rng.randrange(0, len(donor) - isz)
        name = "%s_%d" % (tag, k)
        flip = rng.random() < 0.5        # which mate is read1
        ends = []
        for mate, (st0, fwd) in enumerate(((s, True), (s + isz - read_len, False))):
            seq = donor[st0:st0 + read_len]
            pcs = align_read(g, segs, st0, read_len, rng)
            # primary piece = the longest; everything else is soft clipped
            main = max(pcs, key=lambda p: p[2])
            c, rp, n, strand, qb, qe = main
            lclip, rclip = qb, read_len - qe
            # orientation of the stored read relative to the reference
            stored = seq if strand == "+" else rc(seq)
            if strand == "-":
                lclip, rclip = rclip, lclip
            read_rev = (not fwd) ^ (strand == "-")   # mate2 is sequenced from the reverse strand
            ends.append(dict(name=name, contig=c, pos=rp, reflen=n, lclip=lclip, rclip=rclip, seq=stored,
                             rev=read_rev, pieces=pcs, main=main, donor_fwd=fwd))
        r1, r2 = (ends[1], ends[0]) if flip else (ends[0], ends[1])
        r1["first"], r2["first"] = True, False
        for a, b in ((r1, r2), (r2, r1)):
            a["mate"] = b
        recs.extend((r1, r2))
    return recs


def phred_string(rng: random.Random, n: int) -> str:
    out = []
    for i in range(n):
        r = rng.random()
        if r < 0.80:
            out.append("I")
        elif r < 0.93:
            out.append("H")
        else:
            out.append(chr(33 + rng.randrange(2, 41)))
    # quality tails like the bundled example
    for i in range(1, min(5, n)):
        if rng.random() < 0.5:
            out[-i] = "D"
    return "".join(out)


def to_sam(g: Genome, recs: List[dict], rng: random.Random, opts) -> List[tuple]:
    """Returns list of (tid, pos0, line) records, unsorted."""
    names = g.names
    out = []
    for r in recs:
        r["dup"] = False
    # mark some pairs as duplicates, unmapped mates, low mapq
    for r in recs:
        if not r["first"]:
            continue
        m = r["mate"]
        x = rng.random()
        if x < opts.get("p_dup", 0.01):
            r["dup"] = m["dup"] = True
        elif x < opts.get("p_dup", 0.01) + opts.get("p_unmapped_mate", 0.01):
            m["unmapped"] = True
        elif x < 0.03 + opts.get("p_lowq", 0.01):
            r["mapq"] = rng.choice((0, 0, 5, 15))
    for r in recs:
        m = r["mate"]
        L = len(r["seq"])
        unmapped = r.get("unmapped", False)
        m_unmapped = m.get("unmapped", False)
        seq = r["seq"]
        qual = phred_string(rng, L)
        lclip, rclip, n = r["lclip"], r["rclip"], r["reflen"]
        # background random soft clip on fully aligned reads
        if lclip == 0 and rclip == 0 and rng.random() < opts.get("p_bgclip", 0.01):
            k = rng.randrange(3, 31)
            if rng.random() < 0.5:
                lclip = k
                seq = "".join(rng.choice("ACGT") for _ in range(k)) + seq[k:]
                r["pos"] += k
            else:
                rclip = k
                seq = seq[:L - k] + "".join(rng.choice("ACGT") for _ in range(k))
            n -= k
        # sequencing errors inside the aligned part
        sl = list(seq)
        for i in range(L):
            if rng.random() < opts.get("p_err", 0.002):
                sl[i] = rng.choice("ACGT")
        seq = "".join(sl)
        # CIGAR for the aligned part, with rare I / D / N / = / X decoration
        mid = []
        x = rng.random()
        reflen = n
        if n > 40 and x < opts.get("p_indel", 0.01):
            a = rng.randrange(10, n - 20)
            if rng.random() < 0.5:
                k = rng.randrange(1, 4)       # insertion: consumes query only
                mid = [(a, "M"), (k, "I"), (n - a - k, "M")]
                reflen = n - k
            else:
                k = rng.randrange(1, 6)       # deletion (or N): consumes reference only
                mid = [(a, "M"), (k, "N" if rng.random() < 0.3 else "D"), (n - a, "M")]
                reflen = n + k
        elif n > 40 and x < opts.get("p_indel", 0.01) + opts.get("p_eqx", 0.0):
            a = rng.randrange(5, n - 10)
            mid = [(a, "="), (1, "X"), (n - a - 1, "=")]
        else:
            mid = [(n, "M")]
        cig = ([(lclip, "S")] if lclip else []) + mid + ([(rclip, "S")] if rclip else [])
        flag = 1
        flag |= 64 if r["first"] else 128
        if r["rev"]:
            flag |= 16
        if m["rev"]:
            flag |= 32
        if r["dup"]:
            flag |= 1024
        mapq = r.get("mapq", 60)
        tid = names.index(r["contig"])
        pos = r["pos"]
        mtid = names.index(m["contig"])
        mpos = m["pos"]
        if unmapped:
            flag |= 4
            flag &= ~16
            tid, pos = mtid, mpos
            cig, mapq = [], 0
            seq_out = seq if not r["rev"] else rc(seq)
            qual_out = qual
        else:
            seq_out, qual_out = seq, qual
        if m_unmapped:
            flag |= 8
            flag &= ~32
            mtid, mpos = tid, pos
        isize = 0
        if not unmapped and not m_unmapped and tid == mtid:
            left = min(pos, mpos)
            right = max(pos + reflen, mpos + m["reflen"])
            isize = right - left
            if pos > mpos or (pos == mpos and not r["first"]):
                isize = -isize
            # proper pair: FR orientation, sensible size
            fr = (not r["rev"] and m["rev"] and pos <= mpos) or (r["rev"] and not m["rev"] and mpos <= pos)
            if fr and abs(isize) < opts.get("max_proper", 1000):
                flag |= 2
        aux = ""
        if rng.random() < opts.get("p_xc", 0.003) and not unmapped:
            aux = "\tXC:i:%d" % rng.randrange(20, 90)
        if rng.random() < 0.3:
            aux += "\tNM:i:%d\tMD:Z:%d" % (rng.randrange(0, 3), reflen)
        if rng.random() < 0.1:
            aux += "\tXT:A:U"
        cig_s = "".join("%d%s" % p for p in cig) if cig else "*"
        end0 = pos + (reflen if cig else 1)
        line = "\t".join(map(str, (
            r["name"], flag, names[tid], pos + 1, mapq, cig_s,
            "=" if mtid == tid else names[mtid], mpos + 1, isize, seq_out, qual_out))) + aux
        out.append((tid, pos, line))
        # hard-clipped supplementary record for the second-longest piece of a split read
        if len(r["pieces"]) > 1 and not unmapped and rng.random() < opts.get("p_supp", 0.3):
            others = [p for p in r["pieces"] if p is not r["main"]]
            c2, rp2, n2, st2, qb2, qe2 = max(others, key=lambda p: p[2])
            if n2 >= 20:
                full = r["seq"] if r["main"][3] == "+" else rc(r["seq"])     # donor-forward read
                piece = full[qb2:qe2]
                hl, hr = qb2, L - qe2
                if st2 == "-":
                    piece = rc(piece)
                    hl, hr = hr, hl
                sflag = (flag & ~(2 | 16)) | 2048
                if (not r["donor_fwd"]) ^ (st2 == "-"):
                    sflag |= 16
                cg = ("%dH" % hl if hl else "") + "%dM" % n2 + ("%dH" % hr if hr else "")
                t2 = names.index(c2)
                sline = "\t".join(map(str, (
                    r["name"], sflag, c2, rp2 + 1, 60, cg, "=" if mtid == t2 else names[mtid], mpos + 1, 0, piece,
                    phred_string(rng, n2))))
                out.append((t2, rp2, sline))
    return out


def build_sample(seed: int, contigs, donor_chroms, coverage_ref: float, coverage_donor: float, read_len=100,
                 isize_mu=400, isize_sd=30, opts=None):
    """Returns (genome, header_lines, sorted sam lines)."""
    opts = opts or {}
    rng = random.Random(seed)
    g = Genome(random.Random(opts.get("genome_seed", 12345)), contigs)
    recs = []
    for c, n in contigs:
        recs += make_records(g, [(c, 0, n, "+")], rng, coverage_ref * opts.get("cov_scale", {}).get(c, 1.0),
                             read_len, isize_mu, isize_sd, "r%s" % c, opts)
    for i, segs in enumerate(donor_chroms):
        recs += make_records(g, segs, rng, coverage_donor, read_len, isize_mu, isize_sd, "d%d" % i, opts)
    lines = to_sam(g, recs, rng, opts)
    lines.sort(key=lambda t: (t[0], t[1]))
    header = ["@HD\tVN:1.0\tSO:coordinate"] + ["@SQ\tSN:%s\tLN:%d" % (c, n) for c, n in contigs]
    return g, header, [l for _, _, l in lines]


MICRO_CONTIGS = [("chr2", 24000), ("chr10", 16000), ("HBV", 3215)]


def micro_donors():
    """deletion + inversion + tandem duplication on chr2, a chr2->chr10 translocation, an HBV integration
    into chr10."""
    d1 = [("chr2", 0, 3000, "+"), ("chr2", 3400, 7000, "+"),          # DEL 3000-3400
          ("chr2", 7000, 8200, "-"),                                    # INV 7000-8200
          ("chr2", 8200, 12000, "+"), ("chr2", 11500, 12000, "+"),     # tandem DUP 11500-12000
          ("chr2", 12000, 15000, "+"), ("chr10", 9000, 16000, "+")]    # CTX chr2:15000 -> chr10:9000
    d2 = [("chr10", 0, 4000, "+"), ("HBV", 500, 2500, "+"), ("chr10", 4000, 9000, "+"),   # virus integration
          ("chr2", 15000, 24000, "+")]
    return [d1, d2]


def micro_sample(kind: str):
    """kind: 'tumor' (all SVs, 40x + 40x) or 'normal' (only the deletion germline, lower depth)"""
    if kind == "tumor":
        return build_sample(20261017, MICRO_CONTIGS, micro_donors(), 20, 20,
                            opts=dict(cov_scale={"HBV": 3.0}))
    germ = [[("chr2", 0, 3000, "+"), ("chr2", 3400, 24000, "+")]]
    return build_sample(20261018, MICRO_CONTIGS, germ, 15, 12, opts=dict(cov_scale={"HBV": 0.2}))


def write_fasta(g: Genome, path: str):
    with open(path, "w") as f:
        for c in g.names:
            f.write(">%s\n" % c)
            s = g.seq[c]
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")
