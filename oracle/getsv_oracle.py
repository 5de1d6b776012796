"""CPU restatement of `seeksv getsv` and `seeksv somatic` (reference source v1.2.3) - TEST
INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this module.
Pure-Python loops (numpy only for the depth arrays): sized for inputs of up to ~10^5 records.

Pinned against the reference binary's own output (oracle/_ref/seeksv == prebuilt v1.2.3) for
example/{cancer,normal} (sv files, stdout, somatic) and the hand-written cases in tests/golden/
(tests/test_oracle_golden.py). NOT restated: libbam's pileup read cap (maxcnt = 8000 live reads,
SURVEY.md quirk Q12) - depth at positions covered by more than 8000 reads is "parity unpinned".

Citations are relative to /root/reference/seeksv/.
"""
from __future__ import annotations

import bisect
import math
from typing import Dict, List, Optional, Tuple

import numpy as np

from .bamio import (CIGAR_OPS, FDUP, FMREVERSE, FMUNMAP, FPAIRED, FPROPER, FREVERSE, FSECONDARY, FUNMAP,
                    Header, Rec)
from .getclip_oracle import (OP_D, OP_EQ, OP_H, OP_M, OP_N, OP_S, OP_X, cigar_text, generate_cigar,
                             match_begin_first, match_end_first)

K_CROSS = 5   # kCrossLength, getsv.cpp:15


def fmt_double(x: float) -> str:
    """ostream << double with default precision 6 (getsv.cpp:1848,1857)."""
    return "%g" % x


def is_hard_clip(rec: Rec) -> bool:
    """IsHardClip, clip_reads.cpp:247-257 (an empty CIGAR is treated as "not hard clipped"; the
    reference reads adjacent bytes there, which never decode to op H for the inputs of this path)."""
    return bool(rec.cigar) and (rec.cigar[0][1] == OP_H or rec.cigar[-1][1] == OP_H)


def revcomp(s: str) -> str:
    """GetReverseComplementSeq, clip_reads.cpp:414-466: A<->T, C<->G, case folded to upper for
    acgtn, every other character kept."""
    m = {"A": "T", "a": "T", "T": "A", "t": "A", "C": "G", "c": "G", "G": "C", "g": "C", "N": "N", "n": "N"}
    return "".join(m.get(c, c) for c in reversed(s))


def change_cigar_type(s: str) -> List[Tuple[int, str]]:
    """ChangeCigarType, getsv.cpp:433-451."""
    out, n = [], 0
    for ch in s:
        if ch.isdigit():
            n = n * 10 + ord(ch) - 48
        else:
            out.append((n, ch))
            n = 0
    return out


def calend(rec: Rec) -> int:
    """bam_calend of the LINKED libbam (0-based exclusive end): pos + sum of M, D, N lengths only.
    sam/bam.h:650 declares it; the archive is samtools-0.1.16-era code in which `=` and `X` do not
    advance the reference (probed: `oracle/_ref/bamtool calend 100 10M5X` -> 110, `10M5=` -> 110,
    `10M5D` / `10M5N` -> 115; pinned in tests/test_oracle_golden.py)."""
    e = rec.pos
    for ln, op in rec.cigar:
        if op in (OP_M, OP_D, OP_N):
            e += ln
    return e


# ----------------------------------------------------------------------------------------------
# join of clip.gz with clip.bam/sam -> junctions
# ----------------------------------------------------------------------------------------------
class AlignInfo:
    """AlignInfo, getsv.h:24-45."""
    __slots__ = ("chr", "pos", "len", "strand", "cigar", "seq", "lclip", "rclip", "type")


def get_align_info(h: Header, b: Rec) -> AlignInfo:
    """GetAlignInfo, getsv.cpp:25-71."""
    a = AlignInfo()
    if b.flag & FUNMAP:
        a.chr, a.pos, a.len, a.strand, a.cigar, a.seq, a.lclip, a.rclip, a.type = (
            "Exogenous", -1, -1, "*", [], "", 0, 0, "n")
        return a
    a.type = "r" if (b.flag & FSECONDARY or b.mapq == 0) else "u"
    a.lclip = a.rclip = 0
    if b.cigar:
        if b.cigar[0][1] in (OP_S, OP_H):
            a.lclip = b.cigar[0][0]
        if b.cigar[-1][1] in (OP_S, OP_H):
            a.rclip = b.cigar[-1][0]
    a.cigar, a.len = generate_cigar(b)
    a.strand = "-" if b.flag & FREVERSE else "+"
    a.seq = b.qname      # (reverse-complemented for '-' in the reference; never read afterwards)
    a.chr = h.names[b.tid]
    a.pos = b.pos + 1
    return a


class SeqInfo:
    """SeqInfo, getsv.h:48-70."""
    __slots__ = ("seq", "cigar", "lclip", "rclip", "support", "uniq")

    def __init__(self, seq, cigar, lclip, rclip, support, uniq):
        self.seq, self.cigar, self.lclip, self.rclip, self.support, self.uniq = seq, list(cigar), lclip, rclip, support, uniq


class Other:
    """OtherInfo, getsv.h:88-107."""
    __slots__ = ("up", "down", "micro", "pairs")

    def __init__(self, up, down, micro, pairs):
        self.up, self.down, self.micro, self.pairs = up, down, micro, pairs


def jkey(up_chr, up_pos, up_strand, down_chr, down_pos, down_strand):
    """Junction::operator<, getsv.h:187-225: (up_chr, down_chr, up_strand, down_strand, up_pos, down_pos)."""
    return (up_chr, down_chr, up_strand, down_strand, up_pos, down_pos)


class JunctionMap:
    """std::multimap<Junction, OtherInfo>: sorted by key, equal keys in insertion order."""

    def __init__(self):
        self.keys: List[tuple] = []
        self.vals: List[Other] = []

    def insert(self, key, val):
        i = bisect.bisect_right(self.keys, key)
        self.keys.insert(i, key)
        self.vals.insert(i, val)

    def equal_range(self, key):
        return bisect.bisect_left(self.keys, key), bisect.bisect_right(self.keys, key)

    def erase(self, i):
        del self.keys[i]
        del self.vals[i]

    def __len__(self):
        return len(self.keys)


def get_junction(chr_, pos, cigar_vec, aligned_seq, clipped_seq, support, orientation, ai: AlignInfo,
                 jmap: JunctionMap):
    """GetJunction, getsv.cpp:1705-1845. (type 'n' returns before anything is stored, quirk Q7.)"""
    if ai.type == "u":
        uniq = 2
    elif ai.type == "r":
        uniq = 1
    else:
        return
    cigar_vec = list(cigar_vec)
    if ai.strand == "+":
        if orientation == "5":
            key = jkey(ai.chr, ai.pos + ai.len - 1, "+", chr_, pos, "+")
            up = SeqInfo(clipped_seq, ai.cigar, ai.lclip, ai.rclip, 0, uniq)
            down = SeqInfo(aligned_seq, cigar_vec, 0, 0, support, 0)
        elif orientation == "3":
            key = jkey(chr_, pos, "+", ai.chr, ai.pos, "+")
            up = SeqInfo(aligned_seq, cigar_vec, 0, 0, support, 0)
            down = SeqInfo(clipped_seq, ai.cigar, ai.lclip, ai.rclip, 0, uniq)
        else:
            return
    elif ai.strand == "-":
        if orientation == "5":
            if (ai.chr, ai.pos) <= (chr_, pos):
                key = jkey(ai.chr, ai.pos, "-", chr_, pos, "+")
                up = SeqInfo(clipped_seq, ai.cigar, ai.lclip, ai.rclip, 0, uniq)
                down = SeqInfo(aligned_seq, cigar_vec, 0, 0, support, 0)
            else:
                key = jkey(chr_, pos, "-", ai.chr, ai.pos, "+")
                ai.cigar = ai.cigar[::-1]       # ReverseCigar mutates the stored alignment in place
                up = SeqInfo(revcomp(aligned_seq), cigar_vec[::-1], 0, 0, support, 0)
                down = SeqInfo(revcomp(clipped_seq), ai.cigar, ai.rclip, ai.lclip, 0, uniq)
        elif orientation == "3":
            end = ai.pos + ai.len - 1
            if (chr_, pos) <= (ai.chr, end):
                key = jkey(chr_, pos, "+", ai.chr, end, "-")
                up = SeqInfo(aligned_seq, cigar_vec, 0, 0, support, 0)
                down = SeqInfo(clipped_seq, ai.cigar, ai.lclip, ai.rclip, 0, uniq)
            else:
                key = jkey(ai.chr, end, "+", chr_, pos, "-")
                ai.cigar = ai.cigar[::-1]
                up = SeqInfo(revcomp(clipped_seq), ai.cigar, ai.rclip, ai.lclip, 0, uniq)
                down = SeqInfo(revcomp(aligned_seq), cigar_vec[::-1], 0, 0, support, 0)
        else:
            return
    else:
        return
    lo, hi = jmap.equal_range(key)
    status = True
    for i in range(lo, hi):
        o = jmap.vals[i]
        # getsv.cpp:1817 - clip-length signature; every matching entry accumulates (quirk Q8)
        if o.up.rclip == down.lclip and o.down.lclip == up.rclip:
            o.up.uniq = max(o.up.uniq, up.uniq)
            o.down.uniq = max(o.down.uniq, down.uniq)
            o.up.support += up.support
            o.down.support += down.support
            if o.micro == -1:
                o.micro = 0          # existing.up_pos - junction.up_pos, equal keys -> 0
            status = False
    if status:
        jmap.insert(key, Other(up, down, -1, 0))


def parse_clip_text(text: str):
    """clip.gz lines as read by `fin >> chr >> pos >> orientation >> cigar >> ...` (getsv.h:453-456):
    whitespace-separated tokens, rest of line dropped."""
    out = []
    for line in text.split("\n"):
        t = line.split()
        if len(t) < 9:
            continue
        out.append((t[0], int(t[1]), t[2][0], t[3], t[4], t[5], t[6], t[7], int(t[8])))
    return out


def join_clip_alignments(clip_lines, h: Header, alns: List[Rec], jmap: JunctionMap):
    """InputSoftInfoStoreBreakpoint<T>, getsv.h:423-541, including its quirks: only the
    first line of a run of equal clipped sequences is crossed with the alignments; the first
    alignment of a new run is filed under the previous run's sequence; the trailing loop does not
    skip hard-clipped alignments."""
    it = iter(alns)
    group: List[tuple] = []
    aligns: Dict[tuple, AlignInfo] = {}
    last = ""

    def cross():
        if group:
            chr_, pos, ori, cigar, aseq, aqual, cseq, cqual, sup = group[0]
            cv = change_cigar_type(cigar)
            for k in sorted(aligns):
                get_junction(chr_, pos, cv, aseq, cseq, sup, ori, aligns[k], jmap)

    for line in clip_lines:
        cseq = line[6]
        if last == "" or last == cseq:
            group.append(line)
            last = cseq
            continue
        for b in it:
            if is_hard_clip(b):
                continue
            ai = get_align_info(h, b)
            if last == b.qname:
                aligns.setdefault((last, (ai.chr, ai.pos)), ai)
            else:
                cross()
                group = [line]
                aligns = {(last, (ai.chr, ai.pos)): ai}
                last = cseq
                break
        # alignment stream exhausted: the line is dropped and the state is left as is
    for b in it:
        ai = get_align_info(h, b)
        if last == b.qname:
            aligns.setdefault((last, (ai.chr, ai.pos)), ai)
        else:
            break
    cross()


def merge_junction(jmap: JunctionMap, search_length: int):
    """MergeJunction, getsv.cpp:1325-1482."""
    i = 0
    K, V = jmap.keys, jmap.vals
    while i < len(K):
        a = V[i]
        if a.up.rclip > 0 or a.up.lclip > 0:
            i += 1
            continue
        j = i + 1
        mark = False
        while (j < len(K) and K[i][0] == K[j][0] and K[i][1] == K[j][1] and K[i][2] == K[j][2]
               and K[i][3] == K[j][3] and K[j][4] - K[i][4] <= search_length):
            b = V[j]
            up_strand = K[i][2]
            if abs(K[j][5] - K[i][5]) <= search_length and b.down.lclip == 0:
                u1 = d1 = u2 = d2 = ""
                if len(a.up.cigar) == 1 and len(b.up.cigar) == 1:
                    mh = K[j][4] - K[i][4]
                    if (up_strand == "+" and len(b.up.seq) < mh + 5) or (up_strand == "-" and len(a.up.seq) < mh + 5):
                        j += 1
                        continue
                    if up_strand == "+":
                        u1, d1 = a.up.seq, a.down.seq
                        u2 = b.up.seq[:len(b.up.seq) - mh]
                        d2 = b.up.seq[len(b.up.seq) - mh:] + b.down.seq
                    else:
                        u1 = a.up.seq[:len(a.up.seq) - mh]
                        d1 = a.up.seq[len(a.up.seq) - mh:] + a.down.seq
                        u2, d2 = b.up.seq, b.down.seq
                elif len(a.down.cigar) == 1 and len(b.down.cigar) == 1:
                    mh = abs(K[j][5] - K[i][5])
                    if (up_strand == "+" and len(a.down.seq) < mh + 5) or (up_strand == "-" and len(b.down.seq) < mh + 5):
                        j += 1
                        continue
                    if up_strand == "+":
                        d1 = a.down.seq[mh:]
                        d2 = b.down.seq
                        u1 = a.up.seq + a.down.seq[:mh]
                        u2 = b.up.seq
                    else:
                        d1 = a.down.seq
                        d2 = b.down.seq[mh:]
                        u1 = a.up.seq
                        u2 = b.up.seq + b.down.seq[:mh]
                if match_end_first(u1, u2) >= 0.85 and match_begin_first(d1, d2) >= 0.85:
                    a.up.uniq = max(a.up.uniq, b.up.uniq)
                    a.down.uniq = max(a.down.uniq, b.down.uniq)
                    if a.micro == -1 and b.micro == -1:
                        a.up.support += b.up.support
                        a.down.support += b.down.support
                        if (a.up.support != 0 and b.down.support != 0) or (a.down.support != 0 and b.up.support != 0):
                            a.micro = K[j][4] - K[i][4]
                        jmap.erase(j)
                    elif a.micro != -1 and b.micro == -1:
                        a.up.support += b.up.support
                        a.down.support += b.down.support
                        jmap.erase(j)
                    elif a.micro == -1 and b.micro != -1:
                        b.up.support += a.up.support
                        b.down.support += a.down.support
                        mark = True
                    else:
                        if a.up.support > b.up.support or a.down.support == b.down.support:
                            a.up.support += b.up.support
                            jmap.erase(j)
                        elif a.up.support == b.up.support or a.down.support > b.down.support:
                            a.down.support += b.down.support
                            jmap.erase(j)
                        elif b.up.support > a.up.support and a.down.support == b.down.support:
                            b.up.support += a.up.support
                            mark = True
                        elif b.down.support > a.down.support and b.up.support == a.up.support:
                            b.down.support += a.down.support
                            mark = True
                        else:
                            j += 1
                    if mark:
                        break
                else:
                    j += 1
            else:
                j += 1
        if mark:
            jmap.erase(i)
        else:
            i += 1


# ----------------------------------------------------------------------------------------------
# whole-BAM statistics
# ----------------------------------------------------------------------------------------------
def insert_size_stats(recs: List[Rec], min_mapq: int, pairs_used: int) -> Optional[Tuple[int, int]]:
    """CalculateInsertsizeDeviation, cluster.cpp:15-83. Returns (mean, deviation) or None when no
    pair qualifies (the reference then leaves both at 0)."""
    total, n, sizes = 0, 0, []
    for b in recs:
        if b.mapq < min_mapq:
            continue
        if is_hard_clip(b):
            continue
        if (b.flag & FPAIRED) and (b.flag & FPROPER) and not (b.flag & FDUP) and b.isize > 0:
            total += b.isize
            sizes.append(b.isize)
            n += 1
        if n == pairs_used:
            break
    if n == 0:
        return None
    mean = total // n

    def i32(x):
        return ((x + 2 ** 31) % 2 ** 32) - 2 ** 31
    dev = 0.0
    for s in sizes:
        d = i32(s - mean)
        dev += float(i32(d * d))
    return mean, int(math.sqrt(dev / n))


def is_concordant(b: Rec, mean: int, dev: int, times: int) -> bool:
    """IsConcordant, cluster.cpp:136-147."""
    lo, hi, isz = mean - dev * times, mean + dev * times, b.isize
    if not (b.flag & FREVERSE) and (b.flag & FMREVERSE) and lo <= isz <= hi:
        return True
    if (b.flag & FREVERSE) and not (b.flag & FMREVERSE) and isz < 0:
        return lo <= abs(isz) <= hi
    return False


def discordant_pairs(h: Header, recs: List[Rec], key, min_mapq: int, mean: int, dev: int, times: int) -> int:
    """FindDiscordantReadPairs for one junction, getsv.cpp:1123-1247 (== the body of the all-junction
    overload, getsv.cpp:1039-1117). The bam_iter_query window is restated as "records on tid with
    pos < end and calend > beg" (sam/bam.h:650-670; empty CIGAR -> calend = pos + 1)."""
    up_chr, down_chr, up_strand, down_strand, up_pos, down_pos = key
    lo_is, hi_is = mean - dev * times, mean + dev * times
    if lo_is < 0:
        lo_is = 0
    if up_chr not in h.names:
        return 0
    tid = h.names.index(up_chr)
    chr_len = h.lengths[tid]
    if up_strand == "+":
        end = up_pos
        beg = end - hi_is
    elif up_strand == "-":
        beg = up_pos - 1 - K_CROSS
        end = up_pos - 1 + hi_is
    else:
        return 0
    if beg <= 0:
        beg = 1
    if (end & 0xFFFFFFFF) > chr_len:    # int vs unsigned comparison, getsv.cpp:1060
        end = chr_len
    mtid = h.names.index(down_chr) if down_chr in h.names else -1
    n = 0
    for b in recs:
        if b.tid != tid:
            continue
        rend = calend(b) if b.cigar else b.pos + 1
        if not (b.pos < end and rend > beg):
            continue
        if b.mapq < min_mapq or is_hard_clip(b):
            continue
        if (b.flag & (FDUP | FUNMAP | FMUNMAP)) or is_concordant(b, mean, dev, times):
            continue
        if mtid == -1 or mtid != b.mtid:
            continue
        L = b.l_qseq
        rev, mrev = bool(b.flag & FREVERSE), bool(b.flag & FMREVERSE)
        if (up_strand == "+" and down_strand == "+" and b.pos + L <= up_pos + K_CROSS
                and b.mpos + 1 >= down_pos - K_CROSS):
            if not rev and mrev:
                isz = up_pos - b.pos + b.mpos + L - down_pos + 1
                if tid == mtid and up_pos > down_pos and up_pos - down_pos + 1 + 2 * L <= hi_is:
                    ok = False
                    while isz <= hi_is:
                        if isz >= lo_is:
                            ok = True
                            break
                        isz += up_pos - down_pos + 1
                    n += ok
                elif lo_is <= isz <= hi_is:
                    n += 1
        elif up_strand == "-" and down_strand == "+" and rev and mrev and b.mpos + 1 >= down_pos - K_CROSS:
            isz = b.pos + 1 - up_pos + 1 + b.mpos + L - down_pos + 1
            n += lo_is <= isz <= hi_is
        elif (up_strand == "+" and down_strand == "-" and not rev and not mrev
              and b.pos + L <= up_pos + K_CROSS and b.mpos + L <= down_pos + K_CROSS):
            isz = up_pos - b.pos + down_pos - (b.mpos + L) + 1
            n += lo_is <= isz <= hi_is
    return n


def u32(x: int) -> int:
    return x & 0xFFFFFFFF


def get_break(jmap: JunctionMap, flank: int):
    """GetBreak, getsv.cpp:752-789. Ranges are (chr, begin, end) with UNSIGNED 32-bit begin/end
    (ChrRange, getsv.h:231-258; quirk Q11). Returns (positions set, sorted unique ranges,
    junction key -> 4 ranges)."""
    positions = set()
    ranges = set()
    j2r = {}
    for k in jmap.keys:
        up_chr, down_chr, us, ds, up_pos, down_pos = k
        positions.add((up_chr, up_pos))
        positions.add((down_chr, down_pos))
        if up_chr == down_chr and us == ds:
            d = abs(down_pos - 1 - up_pos)
            l = d if d < flank else flank
        else:
            l = flank
        r1 = (up_chr, u32(up_pos - l + 1), u32(up_pos))
        r2 = (up_chr, u32(up_pos + 1), u32(up_pos + l))
        r3 = (down_chr, u32(down_pos - l), u32(down_pos - 1))
        r4 = (down_chr, u32(down_pos), u32(down_pos + l - 1))
        ranges.update((r1, r2, r3, r4))
        j2r.setdefault(k, (r1, r2, r3, r4))      # map::insert keeps the first
    return positions, sorted(ranges), j2r


def i32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def merge_overlap(ranges) -> Dict[Tuple[str, int], int]:
    """MergeOverlap, getsv.cpp:804-835: begin2end keyed (chr, (int)begin) -> (int)end; map::insert
    keeps the first value for a repeated key."""
    out: Dict[Tuple[str, int], int] = {}
    if not ranges:
        return out            # (the reference inserts an uninitialised window here; harmless garbage)
    chr_, b, e = ranges[0]
    for c, rb, re_ in ranges[1:]:
        if chr_ == c and b <= rb and u32(e + 1) >= rb:
            if re_ > e:
                e = re_
        else:
            out.setdefault((chr_, i32(b)), i32(e))
            chr_, b, e = c, rb, re_
    out.setdefault((chr_, i32(b)), i32(e))
    return out


PILEUP_MAXCNT = 8000   # bam_plp_init default of the linked libbam


def pileup_kept(recs: List[Rec], min_mapq: int, maxcnt: int = PILEUP_MAXCNT) -> List[Rec]:
    """Which reads enter libbam's pileup (bam_plp_push as driven by bam_mplp_auto, bam2depth.cpp:72-75;
    behaviour probed with `oracle/_ref/bamtool depth`, see tests/test_oracle_golden.py):
      * tid >= 0 and (flag & 0x704) == 0 after read_bam turned mapQ < min_mapq into UNMAP
        (bam2depth.h:29-35);
      * a read is refused only when it starts at the iterator's current (tid, pos) - i.e. at the same
        position as the previously accepted read - while more than `maxcnt` buffer nodes are
        allocated. Allocated nodes = 2 (head sentinel + free tail) + accepted reads on this tid whose
        end is >= the current position (nodes are released lazily, one position behind).
    Sequential by nature; coordinate-sorted input assumed (the pileup aborts otherwise)."""
    import heapq
    kept = []
    it_tid, it_pos = 0, 0
    live: List[int] = []          # min-heap of ends of accepted reads on it_tid
    for b in recs:
        if b.tid < 0 or (b.flag & 0x704) or b.mapq < min_mapq:
            continue
        end = calend(b)
        if b.tid == it_tid and b.pos == it_pos:
            if len(live) + 2 > maxcnt:
                continue
            if end > it_pos:
                heapq.heappush(live, end)
            # (a zero-reference-length read at the current position is copied but never linked)
        else:
            if b.tid != it_tid:
                live = []
            it_tid, it_pos = b.tid, b.pos
            while live and live[0] < it_pos:      # nodes with end <= pos-1 were released
                heapq.heappop(live)
            heapq.heappush(live, end)
        kept.append(b)
    return kept


def depth_arrays(h: Header, recs: List[Rec], min_mapq: int) -> Dict[int, np.ndarray]:
    """Per-position depth as main_depth sees it (bam2depth.cpp:75-96): reads accepted by the pileup
    (pileup_kept); a position counts when the read has an M base there. D and N are in the pileup but
    subtracted (bam2depth.cpp:94); `=` and `X` are IGNORED by this libbam's CIGAR walk - they advance
    neither reference nor count (probed with `bamtool depth`); baseQ = 0 never filters.
    Returned arrays are 1-based: arr[p]."""
    out = {}
    for tid, ln in enumerate(h.lengths):
        out[tid] = np.zeros(ln + 2, dtype=np.int64)
    for b in pileup_kept(recs, min_mapq):
        p = b.pos + 1
        d = out[b.tid]
        for ln, op in b.cigar:
            if op == OP_M:
                lo, hi = max(p, 1), min(p + ln, len(d))
                if lo < hi:
                    d[lo:hi] += 1
                p += ln
            elif op in (OP_D, OP_N):
                p += ln
    return out


def main_depth(h: Header, recs: List[Rec], positions, ranges, begin2end, min_mapq: int):
    """main_depth, bam2depth.cpp:17-142: literal restatement of the two map walks, fed by the depth
    arrays. Returns (pos2depth, range2depth)."""
    dep = depth_arrays(h, recs, min_mapq)
    pos2depth = {p: 0 for p in positions}
    range2depth = {r: 0 for r in ranges}
    wkeys = sorted(begin2end)
    for tid, name in enumerate(h.names):
        d = dep[tid]
        covered = np.nonzero(d > 0)[0]
        for P in covered.tolist():            # P = pos + 1 (1-based)
            i = bisect.bisect_right(wkeys, (name, P))
            if i == 0:
                continue
            wk = wkeys[i - 1]
            if wk[0] != name or P > begin2end[wk]:
                continue
            wbeg = u32(wk[1])
            j = bisect.bisect_right(ranges, (name, u32(P + 1), u32(P + 1)))
            if j == 0:
                continue      # bam2depth.cpp:102: the `continue` also skips the point depth at :123-124
            j -= 1
            while j != 0:
                r = ranges[j]
                if r[0] != name or r[1] < wbeg:
                    break
                if P <= r[2]:
                    range2depth[r] += int(d[P])
                j -= 1
            if j == 0:
                r = ranges[0]
                if r[0] == name and r[1] >= wbeg and P <= r[2]:
                    range2depth[r] += int(d[P])
            if (name, P) in pos2depth:
                pos2depth[(name, P)] = int(d[P])
    return pos2depth, range2depth


def sv_type(up_chr, up_pos, us, down_chr, down_pos, ds) -> str:
    """GetSVType, clip_reads.cpp:572-581."""
    if up_chr != down_chr:
        return "CTX"
    if us != ds:
        return "INV"
    if up_pos < down_pos:
        return "DEL"
    if up_pos > down_pos:
        return "INS"
    return "Unknown"


def largest_base_freq(seq: str) -> float:
    """CountLargestBaseFrequency, getsv.cpp:1485-1511."""
    n = len(seq)
    cnt = [0] * 5
    for c in seq:
        cnt["ATCG".find(c.upper()) if c.upper() in "ATCG" else 4] += 1
    return max(cnt) / n if n else float("nan")


SV_HEADER = ("@left_chr\tleft_pos\tleft_strand\tleft_clip_read_NO\tright_chr\tright_pos\tright_strand\t"
             "right_clip_read_NO\tmicrohomology_length\tabnormal_readpair_NO\tsvtype\tleft_pos_depth\t"
             "right_pos_depth\taverage_depth_of_left_pos_5end\taverage_depth_of_left_pos_3end\t"
             "average_depth_of_right_pos_5end\taverage_depth_of_right_pos_3end\tleft_pos_clip_percentage\t"
             "right_pos_clip_percentage\tleft_seq_cigar\tright_seq_cigar\tleft_seq\tright_seq\n")   # seeksv.cpp:308


def output_breakpoints(jmap: JunctionMap, pos2depth, range2depth, j2r, min_clip_sum, min_pairs, freq,
                       min_dist, max_micro, min_seq_len, max_indel) -> Tuple[str, str]:
    """OutputBreakpoint, getsv.cpp:838-987 with OutputOneBreakpoint / OutputFilteredBreakpoint
    (getsv.cpp:1846-1862). Returns (sv file body without header, stdout text)."""
    out, filt = [], []
    for k, o in zip(jmap.keys, jmap.vals):
        up_chr, down_chr, us, ds, up_pos, down_pos = k
        updepth = pos2depth[(up_chr, up_pos)] + o.down.support if (up_chr, up_pos) in pos2depth else 0
        downdepth = pos2depth[(down_chr, down_pos)] + o.up.support if (down_chr, down_pos) in pos2depth else 0
        jr = o.up.support + o.down.support
        rate1 = 0.0 if updepth == 0 else jr / updepth
        rate2 = 0.0 if downdepth == 0 else jr / downdepth
        head = "%s\t%d\t%s\t%d\t%s\t%d\t%s\t%d\t%d\t%d\t%s\t%d\t%d\t" % (
            up_chr, up_pos, us, o.up.support, down_chr, down_pos, ds, o.down.support, o.micro, o.pairs,
            sv_type(up_chr, up_pos, us, down_chr, down_pos, ds), updepth, downdepth)
        tail = "%s\t%s\t%s\t%s\t%s\t%s\n" % (
            fmt_double(rate1), fmt_double(rate2), cigar_text(o.up.cigar, o.up.lclip, o.up.rclip),
            cigar_text(o.down.cigar, o.down.lclip, o.down.rclip), o.up.seq, o.down.seq)

        def reject(reason):
            filt.append(reason + "\t" + head + tail)

        if not (o.up.uniq + o.down.uniq >= 2 or o.pairs > 0):
            reject("mappingQ_too_low")
            continue
        if up_chr == down_chr and abs(up_pos - down_pos) < min_dist:
            reject("distance_too_near")
            continue
        if o.micro > max_micro:
            reject("microhomology_len_too_long")
            continue
        if o.pairs < min_pairs:
            reject("abnormal_read_pair_no_not_pass")
            continue
        if ((o.up.support > 0 and o.down.support > 0 and rate1 < freq and rate2 < freq)
                or (o.up.support == 0 and rate2 < freq) or (o.down.support == 0 and rate1 < freq)):
            reject("frequency_too_low")
            continue
        if o.up.support + o.down.support < min_clip_sum:
            reject("total_clipped_reads_NO_not_pass")
            continue
        if o.pairs == 0:
            if (len(o.up.seq) < o.up.lclip + o.up.rclip + min_seq_len
                    or len(o.down.seq) < o.down.lclip + o.down.rclip + min_seq_len):
                reject("seq_length_too_short")
                continue
            if len(o.up.cigar) > 2 * max_indel + 1 or len(o.down.cigar) > 2 * max_indel + 1:
                reject("seq_with_too_many_indels")
                continue
            if largest_base_freq(o.up.seq) >= 0.8 or largest_base_freq(o.down.seq) >= 0.8:
                reject("repeat_bases")
                continue
        avg = [0, 0, 0, 0]
        if k in j2r:
            for n, r in enumerate(j2r[k]):
                if r in range2depth:
                    avg[n] = u32(range2depth[r] // u32(r[2] - r[1] + 1))
        out.append(head + "%d\t%d\t%d\t%d\t" % tuple(i32(a) for a in avg) + tail)
    return "".join(out), "".join(filt)


def read_breakpoint(text: str, jmap: "JunctionMap"):
    """ReadBreakpoint, getsv.cpp:1291-1323 (getsv -B): junctions of an earlier output file are put into the map before the
    join. The reference reads with `fin >> token`, i.e. by whitespace-separated tokens, not by lines; a line that starts
    with '@' is skipped to its end; after the 23rd token the rest of the line is dropped. A token that does not convert
    ends the loop (the stream fails)."""
    i, n = 0, len(text)

    def skip_ws(i):
        while i < n and text[i] in " \t\n\r\v\f":
            i += 1
        return i

    def token(i):
        i = skip_ws(i)
        j = i
        while j < n and text[j] not in " \t\n\r\v\f":
            j += 1
        return text[i:j], j

    def rest_of_line(i):
        j = text.find("\n", i)
        return n if j < 0 else j + 1
    while True:
        up_chr, i = token(i)
        if not up_chr:
            return
        if up_chr[0] == "@":
            i = rest_of_line(i)
            continue
        tok = []
        for _ in range(22):
            t, i = token(i)
            tok.append(t)
        i = rest_of_line(i)
        try:
            up_pos, up_strand, up_n = int(tok[0]), tok[1], int(tok[2])
            down_chr, down_pos, down_strand, down_n = tok[3], int(tok[4]), tok[5], int(tok[6])
            micro, pairs = int(tok[7]), int(tok[8])
            for k in range(10, 16):
                int(tok[k])
            float(tok[16]), float(tok[17])
        except ValueError:
            return
        if len(up_strand) != 1 or len(down_strand) != 1:
            return  # (`fin >> char` takes one character: longer tokens desynchronise the reference - not modelled)
        up = SeqInfo(tok[20], change_cigar_type(tok[18]), 0, 0, up_n, 0)
        down = SeqInfo(tok[21], change_cigar_type(tok[19]), 0, 0, down_n, 0)
        jmap.insert(jkey(up_chr, up_pos, up_strand, down_chr, down_pos, down_strand), Other(up, down, micro, pairs))


def minus_cigar_right(vec: List[Tuple[int, str]], length: int) -> List[Tuple[int, str]]:
    """MinusCigarRight, clip_reads.cpp:507-546: cut `length` read bases (M / I) off the right end of the op list."""
    total = sum(ln for ln, op in vec if op in "MI")
    if total <= length:
        return list(vec)
    left = total - length
    out = []
    for ln, op in vec:
        if op in "MI":
            if ln >= left:
                out.append((left, op))
                return out
            left -= ln
        out.append((ln, op))
    return out


def add_cigar_left(vec: List[Tuple[int, str]], length: int) -> List[Tuple[int, str]]:
    """AddCigarLeft, clip_reads.cpp:548-558."""
    vec = list(vec)
    if vec[0][1] == "M":
        vec[0] = (vec[0][0] + length, "M")
    else:
        vec.insert(0, (length, "M"))
    return vec


def find_junction(h: Header, recs: List[Rec], min_mapq: int, jmap: "JunctionMap"):
    """FindJunction, process_bwasw.cpp:5-227 (getsv -F): junctions from "connected read-through reads" - a read that the
    aligner reports as two records with the same name, each soft-clipped on one side. The first record of a name is kept
    (std::map), the second one that fits it (same strand + opposite sides, or opposite strands + same side) makes a junction
    and removes the name; records that do not fit change nothing."""
    from .getclip_oracle import generate_cigar, get_seq
    from .bamio import FDUP, FREVERSE, FUNMAP
    pending = {}
    for r in recs:
        if r.mapq < min_mapq or (r.flag & FUNMAP) or not r.cigar:      # __g_skip_aln (sam/sam_view.h:26-39) with g_min_mapQ = -w
            continue
        op1, op2 = "MIDNSHP=X"[r.cigar[0][1]], "MIDNSHP=X"[r.cigar[-1][1]]
        if op1 == "H" or op2 == "H" or (op1 == "S" and op2 == "S") or (op1 == "M" and op2 == "M") or (r.flag & FDUP):
            continue
        vec, maplen = generate_cigar(r)
        if op1 == "S":
            side, left_len = "5", r.cigar[0][0]
            right_len, pos = r.l_qseq - left_len, r.pos + 1
        else:   # (the reference takes every other shape as clipped on the right)
            side, right_len = "3", r.cigar[-1][0]
            left_len, pos = r.l_qseq - right_len, r.pos + maplen
        strand = "-" if r.flag & FREVERSE else "+"
        left, _, right, _ = get_seq(r, 0, left_len, right_len)
        cur = dict(chr=h.names[r.tid], pos=pos, left=left, right=right, cigar=vec, side=side, strand=strand)
        prev = pending.get(r.qname)
        if prev is None:
            pending[r.qname] = cur
            continue
        same_strand_other_side = prev["strand"] == strand and prev["side"] != side
        other_strand_same_side = prev["strand"] != strand and prev["side"] == side
        if not (same_strand_other_side or other_strand_same_side):
            continue
        if same_strand_other_side:
            up, down = (cur, prev) if prev["side"] == "5" else (prev, cur)
            if len(up["left"]) >= len(down["left"]):
                micro = len(up["left"]) - len(down["left"])
                key = jkey(up["chr"], up["pos"] - micro, "+", down["chr"], down["pos"], "+")
                up_i = SeqInfo(down["left"], minus_cigar_right(up["cigar"], micro), 0, 0, 0, 2)
                down_i = SeqInfo(down["right"], down["cigar"], 0, 0, 1, 2)
            else:
                micro = 0
                key = jkey(up["chr"], up["pos"], "+", down["chr"], down["pos"], "+")
                up_i = SeqInfo(down["left"], up["cigar"], 0, len(down["left"]) - len(up["left"]), 0, 2)
                down_i = SeqInfo(down["right"], down["cigar"], 0, 0, 1, 2)
        else:
            up, down = (prev, cur) if (prev["chr"], prev["pos"]) < (cur["chr"], cur["pos"]) else (cur, prev)
            if side == "5":
                if len(up["right"]) >= len(down["left"]):
                    micro = len(up["right"]) - len(down["left"])
                    key = jkey(up["chr"], up["pos"], "-", down["chr"], down["pos"] + micro, "+")
                    up_i = SeqInfo(revcomp(up["right"]), up["cigar"], 0, 0, 0, 2)
                    down_i = SeqInfo(revcomp(up["left"]), add_cigar_left(down["cigar"], micro), 0, 0, 1, 2)
                else:
                    micro = 0
                    key = jkey(up["chr"], up["pos"], "-", down["chr"], down["pos"], "+")
                    up_i = SeqInfo(down["left"], up["cigar"], 0, len(down["left"]) - len(up["right"]), 0, 2)
                    down_i = SeqInfo(down["right"], down["cigar"], 0, 0, 1, 2)
            else:
                if len(up["left"]) >= len(down["right"]):
                    micro = len(up["left"]) - len(down["right"])
                    key = jkey(up["chr"], up["pos"] - micro, "+", down["chr"], down["pos"], "-")
                    up_i = SeqInfo(revcomp(down["right"]), minus_cigar_right(up["cigar"], micro), 0, 0, 0, 2)
                    down_i = SeqInfo(revcomp(down["left"]), down["cigar"], 0, 0, 1, 2)
                else:
                    micro = 0
                    key = jkey(up["chr"], up["pos"], "+", down["chr"], down["pos"], "-")
                    up_i = SeqInfo(up["left"], up["cigar"], 0, 0, 0, 2)
                    down_i = SeqInfo(up["right"], down["cigar"], len(down["right"]) - len(up["left"]), 0, 1, 2)
        lo, hi = jmap.equal_range(key)
        if lo == hi:
            jmap.insert(key, Other(up_i, down_i, micro, 0))
        else:
            o = jmap.vals[lo]
            if len(o.up.seq) != len(up_i.seq) or len(o.down.seq) != len(down_i.seq):
                o.down.support += 1
        del pending[r.qname]


def getsv(h: Header, recs: List[Rec], clip_text: str, clip_h: Header, clip_alns: List[Rec], *, flank=50,
          min_mapq=20, pairs_used=5000000, min_clip_sum=3, min_dist=50, max_micro=50, times=4, min_pairs=0,
          flank_len=200, min_seq_len=30, max_indel=1, freq=0.1, output_depth=True, seed_text=None, connect=None,
          connect_min_mapq=1) -> Tuple[str, str]:
    """CallGetsv, seeksv.cpp:157-364. seed_text: contents of the -B file; connect: (header, records) of the -F file
    (connect_min_mapq = -w). Returns (out.sv.txt contents, stdout)."""
    jmap = JunctionMap()
    if seed_text is not None:
        read_breakpoint(seed_text, jmap)
    if connect is not None:
        find_junction(connect[0], connect[1], connect_min_mapq, jmap)
    join_clip_alignments(parse_clip_text(clip_text), clip_h, clip_alns, jmap)
    merge_junction(jmap, flank)
    if pairs_used >= 100000:
        st = insert_size_stats(recs, min_mapq, pairs_used)
        mean, dev = st if st else (0, 0)
        for k, o in zip(jmap.keys, jmap.vals):
            o.pairs = discordant_pairs(h, recs, k, min_mapq, mean, dev, times)
    else:
        min_pairs = 0
    pos2depth, range2depth, j2r = {}, {}, {}
    if output_depth:
        positions, ranges, j2r = get_break(jmap, flank_len)
        begin2end = merge_overlap(ranges)
        pos2depth, range2depth = main_depth(h, recs, positions, ranges, begin2end, min_mapq)
    else:
        freq = 0
    body, filt = output_breakpoints(jmap, pos2depth, range2depth, j2r, min_clip_sum, min_pairs, freq, min_dist,
                                    max_micro, min_seq_len, max_indel)
    return SV_HEADER + body, filt


# ----------------------------------------------------------------------------------------------
# somatic
# ----------------------------------------------------------------------------------------------
def compare_shifted(seq1: str, seq2: str, seq3: str, seq4: str, rate: float) -> int:
    """Compare, clip_reads.cpp:333-372 (seq2 = 3'-clipped part, seq4 = 3'-aligned part): find the first
    10 bases of seq2 inside seq4, shift the split point there and compare both sides."""
    if len(seq2) < 10:
        return -1
    pos = seq4.find(seq2[:10])
    if pos < 0:
        return -1
    seq5 = seq3 + seq4[:pos]
    seq6 = seq4[pos:]
    if match_end_first(seq1, seq5) >= rate and match_begin_first(seq2, seq6) >= rate:
        return pos
    return -1


class ClipTable:
    """multimap<pair<string,int>, ReadsInfo> of somatic.h:40-70: sorted by (chr, pos), insertion order
    within a key. Entries are (seq_left, seq_right, support)."""

    def __init__(self):
        self.keys: List[Tuple[str, int]] = []
        self.vals: List[Tuple[str, str, int]] = []

    def insert(self, key, val):
        i = bisect.bisect_right(self.keys, key)
        self.keys.insert(i, key)
        self.vals.insert(i, val)

    def equal(self, key):
        return self.vals[bisect.bisect_left(self.keys, key):bisect.bisect_right(self.keys, key)]

    def window(self, chr_, lo, hi):
        """lower_bound((chr, lo)) then while same chr and pos <= hi."""
        i = bisect.bisect_left(self.keys, (chr_, lo))
        while i < len(self.keys) and self.keys[i][0] == chr_ and self.keys[i][1] <= hi:
            yield self.vals[i]
            i += 1


def read_clip_tables(clip_text: str, min_len: int) -> Tuple[ClipTable, ClipTable]:
    """ReadsClipReads<T>, somatic.h:40-70 -> (3'-clipped table, 5'-clipped table)."""
    t3, t5 = ClipTable(), ClipTable()
    for chr_, pos, ori, cigar, aseq, aqual, cseq, cqual, sup in parse_clip_text(clip_text):
        if len(cseq) < min_len:
            continue
        if ori == "3":
            t3.insert((chr_, pos), (aseq, cseq, sup))
        elif ori == "5":
            t5.insert((chr_, pos), (cseq, aseq, sup))
    return t3, t5


def somatic(h: Header, recs: List[Rec], normal_clip_text: str, tumor_sv_text: str, *, rate=0.9, min_mapq=20,
            offset=30, min_len=10, pairs_used=5000000, times=4) -> str:
    """CallSomatic + ReadTumorFileAndOutputSomaticInfo, seeksv.cpp:366-410, somatic.cpp:14-427."""
    t3, t5 = read_clip_tables(normal_clip_text, min_len)
    mean = dev = 0
    if pairs_used >= 100000:
        st = insert_size_stats(recs, min_mapq, pairs_used)
        if st:
            mean, dev = st
    out = []

    def first_match(entries, begin_seq, end_seq):
        for sl, sr, sup in entries:
            if match_begin_first(begin_seq, sr) >= rate and match_end_first(end_seq, sl) >= rate:
                return sup
        return 0

    for line in tumor_sv_text.split("\n"):
        t = line.split()
        if not t:
            continue
        if t[0][0] == "@":
            # `fin >> up_chr; getline(fin, temp)`: first token + rest of the line
            rest = line[line.index(t[0]) + len(t[0]):]
            out.append(t[0] + rest + "\tleft_clip_read_NO_of_control\tright_clip_read_NO_of_control\t"
                       "abnormal_read_pair_no_of_control\n")
            continue
        up_chr, up_pos, us, up_n, down_chr, down_pos, ds, down_n, micro, pairs, svt = (
            t[0], int(t[1]), t[2][0], int(t[3]), t[4], int(t[5]), t[6][0], int(t[7]), int(t[8]), int(t[9]), t[10])
        d = [int(x) for x in t[11:17]]
        r1, r2 = float(t[17]), float(t[18])
        up_cigar, down_cigar, up_seq, down_seq = t[19], t[20], t[21], t[22]
        key = jkey(up_chr, up_pos, us, down_chr, down_pos, ds)
        nl = nr = 0
        written = True
        always_pairs = False
        if us == "+" and ds == "+":
            if micro != -1:
                nr = first_match(t5.equal((down_chr, down_pos)), down_seq, up_seq)
                if len(down_seq) >= micro:
                    nl = first_match(t3.equal((up_chr, up_pos + micro)), down_seq[micro:], up_seq + down_seq[:micro])
                always_pairs = True          # somatic.cpp:111 queries unconditionally
            elif up_n == 0:
                nr = first_match(t5.equal((down_chr, down_pos)), down_seq, up_seq)
                for sl, sr, sup in t3.window(up_chr, up_pos, up_pos + offset):
                    if compare_shifted(sl, sr, up_seq, down_seq, rate) != -1:
                        nl = sup
                        break
            elif down_n == 0:
                nl = first_match(t3.equal((up_chr, up_pos)), down_seq, up_seq)
                for sl, sr, sup in t5.window(down_chr, down_pos - offset, down_pos):
                    if compare_shifted(up_seq, down_seq, sl, sr, rate) != -1:
                        nr = sup
                        break
            else:
                written = False
        elif us == "+" and ds == "-":
            rc_up, rc_down = revcomp(up_seq), revcomp(down_seq)
            if micro != -1:
                nl = first_match(t3.equal((up_chr, up_pos + micro)), down_seq[micro:], up_seq + down_seq[:micro])
                nr = first_match(t3.equal((down_chr, down_pos)), rc_up, rc_down)
            elif up_n == 0:
                nr = first_match(t3.equal((down_chr, down_pos)), rc_up, rc_down)
                for sl, sr, sup in t3.window(up_chr, up_pos, up_pos + offset):
                    if compare_shifted(sl, sr, up_seq, down_seq, rate) != -1:
                        nl = sup
                        break
            elif down_n == 0:
                nl = first_match(t3.equal((up_chr, up_pos)), down_seq, up_seq)
                for sl, sr, sup in t3.window(down_chr, down_pos, down_pos + offset):
                    if compare_shifted(sl, sr, rc_down, rc_up, rate) != -1:
                        nr = sup
                        break
            else:
                written = False
        elif us == "-" and ds == "+":
            rc_up, rc_down = revcomp(up_seq), revcomp(down_seq)
            if micro != -1:
                nl = first_match(t5.equal((up_chr, up_pos)), rc_up, rc_down)
                nr = first_match(t5.equal((down_chr, down_pos - micro)), up_seq[len(up_seq) - micro:] + down_seq,
                                 up_seq[:len(up_seq) - micro])
            elif up_n == 0:
                nr = first_match(t5.equal((down_chr, down_pos)), down_seq, up_seq)
                for sl, sr, sup in t5.window(up_chr, up_pos - offset, up_pos):
                    if compare_shifted(rc_up, rc_down, sl, sr, rate) != -1:
                        nl = sup
                        break
            elif down_n == 0:
                nl = first_match(t5.equal((up_chr, up_pos)), rc_up, rc_down)
                for sl, sr, sup in t5.window(down_chr, down_pos - offset, down_pos):
                    if compare_shifted(up_seq, down_seq, sl, sr, rate) != -1:
                        nr = sup
                        break
            else:
                written = False
        else:
            written = False
        if not written:
            continue
        npairs = 0
        if always_pairs or mean != 0:
            npairs = discordant_pairs(h, recs, key, min_mapq, mean, dev, times)
        out.append("%s\t%d\t%s\t%d\t%s\t%d\t%s\t%d\t%d\t%d\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%s\t%s\t%s\t%s\t%s\t%s\t%d\t%d\t%d\n" % (
            up_chr, up_pos, us, up_n, down_chr, down_pos, ds, down_n, micro, pairs, svt, d[0], d[1], d[2], d[3], d[4],
            d[5], fmt_double(r1), fmt_double(r2), up_cigar, down_cigar, up_seq, down_seq, nl, nr, npairs))
    return "".join(out)
