"""CPU restatement of `seeksv getclip` (reference source v1.2.3) - TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this module.
Pure-Python loops: use it on inputs of up to ~10^5 records; larger parity runs go through the
compiled reference in oracle/_ref (built by oracle/build_ref.sh from /root/reference).

Pinned against: the decompressed clip.gz / clip.fq.gz / unmapped_*.fq.gz that the reference binary
(oracle/_ref/seeksv, byte-identical to the prebuilt /root/reference/seeksv/seeksv) writes for
example/{cancer,normal}.sort.bam and for the hand-written SAM cases under tests/golden/
(tests/test_oracle_golden.py).

Every function cites the reference lines it follows (paths relative to /root/reference/seeksv/).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

from .bamio import (CIGAR_OPS, FDUP, FMUNMAP, FREAD1, FREVERSE, FUNMAP, Header, Rec, aux_get_int)

OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X = range(9)


def generate_cigar(rec: Rec) -> Tuple[List[Tuple[int, str]], int]:
    """GenerateCigar, clip_reads.cpp:309-329: op list without S/H, and l = sum of M, D, =, N lengths
    (X is NOT counted)."""
    l = 0
    vec = []
    for ln, op in rec.cigar:
        if op in (OP_H, OP_S):
            continue
        if op in (OP_M, OP_D, OP_EQ, OP_N):
            l += ln
        vec.append((ln, CIGAR_OPS[op]))
    return vec, l


def cigar_text(vec: List[Tuple[int, str]], left: int = 0, right: int = 0) -> str:
    """DisplayCigarVector, clip_reads.h:489-505."""
    s = ""
    if left > 0:
        s += "%dS" % left
    s += "".join("%d%s" % (ln, op) for ln, op in vec)
    if right > 0:
        s += "%dS" % right
    return s


def get_seq(rec: Rec, begin: int, left_len: int, right_len: int) -> Tuple[str, str, str, str]:
    """GetSeq, clip_reads.cpp:286-306: (seq_left, qual_left, seq_right, qual_right); qual[0]==0xff -> "*"."""
    sl = rec.seq_str(begin, begin + left_len).upper()
    sr = rec.seq_str(begin + left_len, begin + left_len + right_len).upper()
    if rec.l_qseq and rec.qual[0] == 0xFF:
        ql, qr = "*", "*"
    else:
        ql = rec.qual_str(begin, begin + left_len)
        qr = rec.qual_str(begin + left_len, begin + left_len + right_len)
    return sl, ql, sr, qr


def match_end_first(a: str, b: str) -> float:
    """CompareStringEndFirst, clip_reads.cpp:194-205. len 0 -> 0/0 = NaN (compares false)."""
    n = min(len(a), len(b))
    if n == 0:
        return float("nan")
    m = sum(1 for i in range(1, n + 1) if a[-i] == b[-i])
    return m / n


def match_begin_first(a: str, b: str) -> float:
    """CompareStringBeginFirst, clip_reads.cpp:207-217."""
    n = min(len(a), len(b))
    if n == 0:
        return float("nan")
    m = sum(1 for i in range(n) if a[i] == b[i])
    return m / n


class Cluster:
    """ReadsInfo, clip_reads.h:44-84."""
    __slots__ = ("sl", "ql", "sr", "qr", "cigar", "support")

    def __init__(self, sl, ql, sr, qr, cigar):
        self.sl, self.ql, self.sr, self.qr, self.cigar, self.support = sl, ql, sr, qr, cigar, 1

    def absorb(self, sl, ql, sr, qr, cigar, left_clipped: bool):
        """ReadsInfo::ChangeSeqAndQual, clip_reads.cpp:57-108 (+ support_read_no_increase)."""
        len1, len2 = len(self.sl), len(sl)
        n = min(len1, len2)
        csl, cql = list(self.sl), list(self.ql)
        for i in range(n):
            if cql[len1 - 1 - i] < ql[len2 - 1 - i]:
                cql[len1 - 1 - i] = ql[len2 - 1 - i]
                csl[len1 - 1 - i] = sl[len2 - 1 - i]
        self.sl, self.ql = "".join(csl), "".join(cql)
        if len1 <= len2:
            self.sl = sl[:len2 - n] + self.sl
            self.ql = ql[:len2 - n] + self.ql
            if not left_clipped:          # aa == RIGHT_CLIPPED, clip_reads.cpp:80-83
                self.cigar = cigar
        len1, len2 = len(self.sr), len(sr)
        n = min(len1, len2)
        csr, cqr = list(self.sr), list(self.qr)
        for i in range(n):
            if cqr[i] < qr[i]:
                cqr[i] = qr[i]
                csr[i] = sr[i]
        self.sr, self.qr = "".join(csr), "".join(cqr)
        if len1 < len2:
            self.sr += sr[n:]
            self.qr += qr[n:]
            if left_clipped:              # aa == LEFT_CLIPPED, clip_reads.cpp:102-105
                self.cigar = cigar
        self.support += 1


def insert_seq(table: Dict[int, List[Cluster]], pos, sl, ql, sr, qr, cigar, limit, left_clipped):
    """InsertSeq, clip_reads.cpp:260-283: first cluster at this key (insertion order) whose left
    parts match from the end AND right parts match from the beginning at >= limit absorbs the read."""
    lst = table.setdefault(pos, [])
    for c in lst:
        if match_end_first(sl, c.sl) >= limit and match_begin_first(sr, c.sr) >= limit:
            c.absorb(sl, ql, sr, qr, cigar, left_clipped)
            return
    lst.append(Cluster(sl, ql, sr, qr, cigar))


def get_sclip_reads(rec: Rec, tab_l, tab_r, limit: float, min_mapq: int, save_low_quality: bool):
    """GetSClipReads, clip_reads.cpp:112-192."""
    if not rec.cigar:
        return  # the reference reads cigar[-1] of an empty CIGAR (undefined); no valid input has this
    op1, op2 = rec.cigar[0][1], rec.cigar[-1][1]
    if op1 == OP_H or op2 == OP_H or rec.mapq < min_mapq or (rec.flag & FDUP):
        return
    if (op1 == OP_S) != (op2 == OP_S):
        xc = aux_get_int(rec.aux, b"XC")
        if xc != 0 and not save_low_quality:
            return
        cigar, reflen = generate_cigar(rec)
        if op1 == OP_S:
            ll = rec.cigar[0][0]
            rl = rec.l_qseq - ll
            sl, ql, sr, qr = get_seq(rec, 0, ll, rl)
            insert_seq(tab_l, rec.pos + 1, sl, ql, sr, qr, cigar, limit, True)
        else:
            rl = rec.cigar[-1][0]
            ll = rec.l_qseq - rl
            sl, ql, sr, qr = get_seq(rec, 0, ll, rl)
            insert_seq(tab_r, rec.pos + reflen, sl, ql, sr, qr, cigar, limit, False)
    elif op1 == OP_S and op2 == OP_S:
        ll = rec.cigar[0][0]
        rclip = rec.cigar[-1][0]
        mid = rec.l_qseq - ll - rclip
        cigar, reflen = generate_cigar(rec)
        xc = aux_get_int(rec.aux, b"XC")
        if xc != 0 and not save_low_quality:
            if not (rec.flag & FREVERSE):
                sl, ql, sr, qr = get_seq(rec, 0, ll, mid)
                insert_seq(tab_l, rec.pos + 1, sl, ql, sr, qr, cigar, limit, True)
            else:
                sl, ql, sr, qr = get_seq(rec, ll, mid, rclip)
                insert_seq(tab_r, rec.pos + reflen, sl, ql, sr, qr, cigar, limit, False)
        else:
            sl, ql, sr, qr = get_seq(rec, 0, ll, mid)
            insert_seq(tab_l, rec.pos + 1, sl, ql, sr, qr, cigar, limit, True)
            sl, ql, sr, qr = get_seq(rec, ll, mid, rclip)
            insert_seq(tab_r, rec.pos + reflen, sl, ql, sr, qr, cigar, limit, False)


def flush(chr_name: str, tab_l, tab_r, clip_out: List[str], fq_out: List[str]):
    """DisplaySClipReadsAndClipFq, clip_reads.h:300-345: '5' map then '3' map, by position, equal
    positions in insertion order."""
    for pos in sorted(tab_l):
        for c in tab_l[pos]:
            clip_out.append("%s\t%d\t5\t%s\t%s\t%s\t%s\t%s\t%d\n" % (
                chr_name, pos, cigar_text(c.cigar), c.sr, c.qr, c.sl, c.ql, c.support))
            fq_out.append("@%s\n%s\n+\n%s\n" % (c.sl, c.sl, c.ql))
    for pos in sorted(tab_r):
        for c in tab_r[pos]:
            clip_out.append("%s\t%d\t3\t%s\t%s\t%s\t%s\t%s\t%d\n" % (
                chr_name, pos, cigar_text(c.cigar), c.sl, c.ql, c.sr, c.qr, c.support))
            fq_out.append("@%s\n%s\n+\n%s\n" % (c.sr, c.sr, c.qr))
    tab_l.clear()
    tab_r.clear()


def getclip(header: Header, recs: List[Rec], limit: float = 0.9, min_mapq: int = 1,
            save_low_quality: bool = False) -> Tuple[str, str, str, str]:
    """InputBamOutputReads, clip_reads.h:363-484. Returns the decompressed contents of
    (P.clip.gz, P.clip.fq.gz, P.unmapped_1.fq.gz, P.unmapped_2.fq.gz)."""
    tab_l: Dict[int, List[Cluster]] = {}
    tab_r: Dict[int, List[Cluster]] = {}
    clip_out: List[str] = []
    fq_out: List[str] = []
    un1: List[str] = []
    un2: List[str] = []
    pending: Dict[str, Tuple[str, str, str]] = {}     # qname -> (seq, qual, end)
    last_tid = 0
    # note: each flush holds one chromosome only, so ordering by pos == ordering by (chr, pos)
    for rec in recs:
        if rec.flag & (FUNMAP | FMUNMAP):
            # GetSeqAndQual clip_reads.cpp:375-388 + StoreUnmapSeqAndQual clip_reads.h:172-219
            if rec.l_qseq:
                seq = rec.seq_str()
                qual = "*" if rec.qual[0] == 0xFF else rec.qual_str()
            else:
                seq, qual = "", ""
            end = "1" if rec.flag & FREAD1 else "2"
            st = pending.get(rec.qname)
            if st is not None:
                if end == "1" and st[2] == "2":
                    un1.append("@%s/1\n%s\n+\n%s\n" % (rec.qname, seq, qual))
                    un2.append("@%s/2\n%s\n+\n%s\n" % (rec.qname, st[0], st[1]))
                    del pending[rec.qname]
                elif end == "2" and st[2] == "1":
                    un1.append("@%s/1\n%s\n+\n%s\n" % (rec.qname, st[0], st[1]))
                    un2.append("@%s/2\n%s\n+\n%s\n" % (rec.qname, seq, qual))
                    del pending[rec.qname]
            else:
                pending[rec.qname] = (seq, qual, end)
        elif rec.tid == last_tid:
            get_sclip_reads(rec, tab_l, tab_r, limit, min_mapq, save_low_quality)
        else:
            # chromosome switch: flush, and the current record is NOT processed (quirk Q1)
            flush(header.names[last_tid], tab_l, tab_r, clip_out, fq_out)
            last_tid = rec.tid
    flush(header.names[last_tid] if header.names else "", tab_l, tab_r, clip_out, fq_out)
    return "".join(clip_out), "".join(fq_out), "".join(un1), "".join(un2)
