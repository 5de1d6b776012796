"""BAM / SAM readers for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this module;
the product (`seeksv_b200/`) never does.

What it restates: the record model of the samtools-0.1.x library the reference links as a prebuilt
archive (`sam/libbam.a`; only its headers are vendored, so the wire format is taken from them):
  * BGZF container          sam/bgzf.h:34-60   (gzip members with a BC extra field, <= 64 KiB each)
  * bam1_core_t / bam1_t    sam/bam.h:169-198  (32-byte core, then qname, cigar, 4-bit seq, qual, aux)
  * CIGAR encoding          sam/bam.h:128-151  ("MIDNSHP=X", op in the low 4 bits)
  * 4-bit base table        sam/bam.h:282      ("=ACMGRSVTWYHKDBN")
  * aux tag access          sam/bam.h:556-566  (bam_aux_get / bam_aux2i)
"""
from __future__ import annotations

import gzip
import re
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

CIGAR_OPS = "MIDNSHP=X"
NT16 = "=ACMGRSVTWYHKDBN"

FPAIRED, FPROPER, FUNMAP, FMUNMAP, FREVERSE, FMREVERSE, FREAD1, FREAD2, FSECONDARY, FQCFAIL, FDUP = (
    1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)


@dataclass
class Header:
    names: List[str]
    lengths: List[int]
    text: str = ""


@dataclass
class Rec:
    tid: int
    pos: int          # 0-based
    mapq: int
    bin: int
    flag: int
    l_qseq: int
    mtid: int
    mpos: int
    isize: int
    qname: str
    cigar: List[Tuple[int, int]]   # (len, op index into CIGAR_OPS)
    seq4: bytes                    # packed 4-bit bases
    qual: bytes                    # raw phred, l_qseq bytes
    aux: bytes
    size: int = 0                  # 4 + block_size of the packed record

    def seq_str(self, beg: int = 0, end: Optional[int] = None) -> str:
        end = self.l_qseq if end is None else end
        s = self.seq4
        return "".join(NT16[(s[i >> 1] >> 4) & 15] if (i & 1) == 0 else NT16[s[i >> 1] & 15] for i in range(beg, end))

    def qual_str(self, beg: int = 0, end: Optional[int] = None) -> str:
        end = self.l_qseq if end is None else end
        return "".join(chr(q + 33) for q in self.qual[beg:end])


def read_bgzf(path: str) -> bytes:
    """Whole uncompressed payload of a BGZF (or plain gzip) file. sam/bgzf.h:34-60."""
    with gzip.open(path, "rb") as f:
        return f.read()


def parse_bam_stream(data: bytes) -> Tuple[Header, List[Rec], int]:
    """Parse 'BAM\\1' header + records. Returns (header, records, offset_of_first_record)."""
    assert data[:4] == b"BAM\1", "not a BAM stream"
    l_text = struct.unpack_from("<i", data, 4)[0]
    text = data[8:8 + l_text].decode("latin-1")
    o = 8 + l_text
    n_ref = struct.unpack_from("<i", data, o)[0]
    o += 4
    names, lengths = [], []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, o)[0]
        o += 4
        names.append(data[o:o + l_name - 1].decode("latin-1"))
        o += l_name
        lengths.append(struct.unpack_from("<i", data, o)[0])
        o += 4
    first = o
    recs = []
    n = len(data)
    while o + 4 <= n:
        block_size = struct.unpack_from("<i", data, o)[0]
        tid, pos, l_qname, mapq, bin_, n_cigar, flag, l_qseq, mtid, mpos, isize = struct.unpack_from(
            "<iiBBHHHiiii", data, o + 4)
        p = o + 36
        # bam1_qname is a C string: it ends at the first NUL (normally at l_qname - 1; the records libbam's SAM reader makes
        # from names of 255+ characters carry no NUL there and the name runs on into the CIGAR words)
        nul = data.find(b"\0", p, o + 4 + block_size)
        qname = data[p:nul if nul >= 0 else o + 4 + block_size].decode("latin-1")
        p += l_qname
        cig = struct.unpack_from("<%dI" % n_cigar, data, p) if n_cigar else ()
        cigar = [(c >> 4, c & 15) for c in cig]
        p += 4 * n_cigar
        seq4 = data[p:p + (l_qseq + 1) // 2]
        p += (l_qseq + 1) // 2
        qual = data[p:p + l_qseq]
        p += l_qseq
        aux = data[p:o + 4 + block_size]
        recs.append(Rec(tid, pos, mapq, bin_, flag, l_qseq, mtid, mpos, isize, qname, cigar, seq4, qual, aux,
                        4 + block_size))
        o += 4 + block_size
    return Header(names, lengths, text), recs, first


def read_bam(path: str) -> Tuple[Header, List[Rec]]:
    h, r, _ = parse_bam_stream(read_bgzf(path))
    return h, r


_BASE2NIB = {c: i for i, c in enumerate(NT16)}


def sam_flag(text: str) -> int:
    """FLAG column as the linked libbam's sam_read1 takes it (probed with `bamtool sam2bam`): a number in any C base
    (strtol base 0), else a string of flag letters."""
    if text[:1].isdigit():
        m = re.match(r"0[xX][0-9a-fA-F]+|0[0-7]*|[0-9]+", text)
        t = m.group(0)
        return int(t, 16) if t[:2] in ("0x", "0X") else int(t, 8) if len(t) > 1 and t[0] == "0" else int(t)
    f = 0
    for c in text:
        k = "pPuUrR12sfd".find(c)
        if k >= 0:
            f |= 1 << k
    return f


def _c_int(text: str) -> int:
    m = re.match(r"\s*[+-]?\d+", text)
    return int(m.group(0)) if m else 0


def _c_float(text: str) -> float:
    m = re.match(r"\s*[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", text)
    return float(m.group(0)) if m else 0.0


def sam_line_to_record(line: str, names: List[str]) -> bytes:
    """One SAM line -> the packed BAM record libbam's sam_read1 builds (bam_import.o of sam/libbam.a; call sites
    clip_reads.h:375, getsv.h:445), byte for byte - pinned against `bamtool sam2bam` in tests/test_sam_text.py:
    CIGAR "*" sets the unmapped flag; the bin is computed from pos / end even without a reference; aux integers narrow
    to the smallest type (c for -127..-1, s, i / C, S, I); l_qname is 8 bits wide, so a name of 255+ characters is cut
    to (length + 1) & 0xff bytes WITHOUT its NUL."""
    t = line.split("\t")
    qname, flag_s, rname, pos, mapq, cigar_s, rnext, pnext, tlen, seq, qual = t[:11]
    flag = sam_flag(flag_s)
    tid = names.index(rname) if rname in names and rname != "*" else -1
    if rnext == "=":
        mtid = tid
    else:
        mtid = names.index(rnext) if rnext in names and rnext != "*" else -1
    pos0 = _c_int(pos) - 1
    cigar = []
    end = pos0
    if cigar_s != "*":
        num = 0
        for ch in cigar_s:
            if ch.isdigit():
                num = num * 10 + ord(ch) - 48
            else:
                op = CIGAR_OPS.index(ch)
                cigar.append((num << 4) | op)
                if op in (0, 2, 3):
                    end += num
                num = 0
    else:
        flag |= FUNMAP
    if end == pos0:
        end = pos0 + 1
    if seq == "*":
        l_qseq, seq4, q = 0, b"", b""
    else:
        l_qseq = len(seq)
        nib = [_BASE2NIB.get(c.upper(), 15) for c in seq]
        if l_qseq & 1:
            nib.append(0)
        seq4 = bytes((nib[i] << 4) | nib[i + 1] for i in range(0, len(nib), 2))
        q = bytes([0xFF] * l_qseq) if qual == "*" else bytes((ord(c) - 33) & 0xFF for c in qual)
    aux = b""
    for fld in t[11:]:
        if len(fld) < 5 or fld[2] != ":" or fld[4] != ":":
            continue
        tag, ty, val = fld[:2].encode("latin-1"), fld[3], fld[5:]
        if ty in "iI":
            v = _c_int(val)
            if v < 0:
                aux += tag + (b"c" + struct.pack("<b", v) if v >= -127 else b"s" + struct.pack("<h", v) if v >= -32767
                              else b"i" + struct.pack("<I", v & 0xFFFFFFFF))
            else:
                aux += tag + (b"C" + struct.pack("<B", v) if v <= 255 else b"S" + struct.pack("<H", v) if v <= 65535
                              else b"I" + struct.pack("<I", v & 0xFFFFFFFF))
        elif ty in "ZH":
            aux += tag + ty.encode() + val.encode("latin-1") + b"\0"
        elif ty in "AacC":
            aux += tag + b"A" + val.encode("latin-1")[:1]
        elif ty == "f":
            aux += tag + b"f" + struct.pack("<f", _c_float(val))
        elif ty == "d":
            aux += tag + b"d" + struct.pack("<d", _c_float(val))
        elif ty == "B" and val:
            sub, items = val[0], val.split(",")[1:]
            aux += tag + b"B" + sub.encode() + struct.pack("<i", len(items))
            for it in items:
                if sub == "f":
                    aux += struct.pack("<f", _c_float(it))
                else:
                    v = int(it, 0)
                    aux += struct.pack({"c": "<B", "C": "<B", "s": "<H", "S": "<H"}.get(sub, "<I"),
                                       v & {"c": 0xFF, "C": 0xFF, "s": 0xFFFF, "S": 0xFFFF}.get(sub, 0xFFFFFFFF))
    qn = qname.encode("latin-1")
    l_qname = (len(qn) + 1) & 0xFF
    name_bytes = qn + b"\0" if l_qname == len(qn) + 1 else qn[:l_qname]
    body = struct.pack("<iiBBHHHiiii", tid, pos0, l_qname, _c_int(mapq) & 0xFF, reg2bin(pos0, end), len(cigar), flag & 0xFFFF,
                       l_qseq, mtid, _c_int(pnext) - 1, _c_int(tlen))
    body += name_bytes + b"".join(struct.pack("<I", c) for c in cigar) + seq4 + q + aux
    return struct.pack("<i", len(body)) + body


def sam_to_stream(path: str) -> bytes:
    """SAM text file -> the uncompressed BAM stream samopen(fn, "r") + samwrite of the linked libbam produce."""
    opener = gzip.open if path.endswith(".gz") else open
    names, lengths, text, recs = [], [], [], []
    with opener(path, "rt", encoding="latin-1", newline="") as f:
        for line in f.read().split("\n"):
            line = line.rstrip("\r")
            if not line:
                continue
            if line[0] == "@" and not recs:
                text.append(line + "\n")
                if line.startswith("@SQ"):
                    sn, ln = None, 0
                    for fld in line.split("\t")[1:]:
                        if fld.startswith("SN:"):
                            sn = fld[3:]
                        elif fld.startswith("LN:"):
                            ln = _c_int(fld[3:])
                    names.append(sn)
                    lengths.append(ln)
                continue
            recs.append(sam_line_to_record(line, names))
    return header_bytes(Header(names, lengths, "".join(text))) + b"".join(recs)


def parse_sam(path: str) -> Tuple[Header, List[Rec]]:
    """SAM text -> the same record model, through the exact bytes samopen(fn, "r") yields (used for the clip.sam hand-off
    and for hand-written known-answer inputs)."""
    h, r, _ = parse_bam_stream(sam_to_stream(path))
    return h, r


def read_alignments(path: str) -> Tuple[Header, List[Rec]]:
    """The reference opens a file as BAM iff its name ends in ".bam" (clip_reads.h:367-373)."""
    return read_bam(path) if path.endswith(".bam") else parse_sam(path)


def aux_get_int(aux: bytes, tag: bytes) -> int:
    """bam_aux2i(bam_aux_get(b, tag)) of the libbam the reference links (call sites clip_reads.cpp:126-127,158-159;
    declared sam/bam.h:556-561), pinned by probing the archive (`oracle/_ref/bamtool auxi`, tests/test_oracle_golden.py):

      * the walk upper-cases the type before asking for its size, and the size table only knows 'C'/'A' (1), 'S' (2) and
        'I' (4) in upper case: a float ('f') or double ('d') field is NOT skipped - its value bytes are read as the next
        tag - so an XC behind one is normally not found; inside a 'B' array the raw sub-type is used, where 'f' is known;
      * the value comes back as int32 ('I' above 2^31 wraps); non-integer types give 0.

    Bytes past the end of the record (only reachable after such a mis-step) are taken as 0 here; the library would read
    whatever the previous record left in its buffer."""
    return aux_walk(aux, tag)[0]


def aux_walk(aux: bytes, tag: bytes) -> Tuple[int, bool]:
    """(value, overrun): the walk of aux_get_int; `overrun` says that it read past the end of the record, where the library
    sees stale bytes of earlier records (the fixtures avoid such records: the reference's result is then undefined)."""
    n = len(aux)
    over = [False]

    def at(i):
        if 0 <= i < n:
            return aux[i]
        over[0] = True
        return 0

    s = 0
    while s < n:
        hit = at(s) == tag[0] and at(s + 1) == tag[1]
        s += 2
        if hit:
            ty = chr(at(s))
            s += 1
            if ty == "c":
                v = at(s)
                return (v - 256 if v > 127 else v), over[0]
            if ty == "C":
                return at(s), over[0]
            if ty == "s":
                v = at(s) | at(s + 1) << 8
                return (v - 65536 if v > 32767 else v), over[0]
            if ty == "S":
                return at(s) | at(s + 1) << 8, over[0]
            if ty in "iI":
                v = at(s) | at(s + 1) << 8 | at(s + 2) << 16 | at(s + 3) << 24
                return (v - (1 << 32) if v >= (1 << 31) else v), over[0]
            return 0, over[0]
        u = chr(at(s)).upper()
        s += 1
        if u in "ZH":
            while s < n and aux[s] != 0:
                s += 1
            if s >= n:
                over[0] = True
            s += 1
        elif u == "B":
            sub = chr(at(s))
            cnt = at(s + 1) | at(s + 2) << 8 | at(s + 3) << 16 | at(s + 4) << 24
            if cnt >= (1 << 31):
                cnt -= 1 << 32
            step = (5 + cnt * (1 if sub in "cCA" else 2 if sub in "sS" else 4 if sub in "iIf" else 0)) & 0xFFFFFFFF
            if step >= (1 << 31):  # 32-bit int arithmetic in the library: a wrapped product walks backwards, off the record
                return 0, True
            s += step
        else:
            s += 1 if u in "CA" else 2 if u == "S" else 4 if u == "I" else 0
    return 0, over[0]


# ------------------------------------------------------------------------------------------------
# writers (fixtures only)
# ------------------------------------------------------------------------------------------------
def reg2bin(beg: int, end: int) -> int:
    """bam_reg2bin, sam/bam.h:700-709."""
    end -= 1
    if beg >> 14 == end >> 14:
        return 4681 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return 585 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return 73 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return 9 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return 1 + (beg >> 26)
    return 0


def pack_record(r: Rec) -> bytes:
    qn = r.qname.encode("latin-1") + b"\0"
    end = r.pos
    for ln, op in r.cigar:
        if op in (0, 2, 3):
            end += ln
    if end == r.pos:
        end = r.pos + 1
    b = reg2bin(r.pos, end) if r.tid >= 0 else 4680
    body = struct.pack("<iiBBHHHiiii", r.tid, r.pos, len(qn), r.mapq, b, len(r.cigar), r.flag, r.l_qseq, r.mtid,
                       r.mpos, r.isize)
    body += qn + b"".join(struct.pack("<I", (ln << 4) | op) for ln, op in r.cigar) + r.seq4 + r.qual + r.aux
    return struct.pack("<i", len(body)) + body


def header_bytes(h: Header) -> bytes:
    text = h.text.encode("latin-1")
    out = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(h.names))
    for n, l in zip(h.names, h.lengths):
        nb = n.encode("latin-1") + b"\0"
        out += struct.pack("<i", len(nb)) + nb + struct.pack("<i", l)
    return out


def bgzf_compress(data: bytes, level: int = 1, block: int = 0xff00, strategy: int = 0) -> bytes:
    """BGZF container, sam/bgzf.h:34-60: gzip members with the 'BC' extra sub-field + EOF marker.
    strategy: zlib strategy (0 default, 4 = Z_FIXED forces fixed-Huffman blocks)."""
    import zlib
    out = []
    for i in list(range(0, len(data), block)) + [None]:
        chunk = b"" if i is None else data[i:i + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        comp = c.compress(chunk) + c.flush()
        bsize = len(comp) + 25
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + comp +
                   struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return b"".join(out)


def write_bam(path: str, h: Header, recs: List[Rec], level: int = 1):
    with open(path, "wb") as f:
        f.write(bgzf_compress(header_bytes(h) + b"".join(pack_record(r) for r in recs), level))


def make_rec(qname, flag, tid, pos0, mapq, cigar_s, mtid, mpos0, isize, seq, qual, aux=b"") -> Rec:
    cigar, num = [], 0
    if cigar_s != "*":
        for ch in cigar_s:
            if ch.isdigit():
                num = num * 10 + ord(ch) - 48
            else:
                cigar.append((num, CIGAR_OPS.index(ch)))
                num = 0
    nib = [_BASE2NIB.get(c.upper(), 15) for c in seq]
    if len(nib) & 1:
        nib.append(0)
    seq4 = bytes((nib[i] << 4) | nib[i + 1] for i in range(0, len(nib), 2))
    q = bytes([0xFF] * len(seq)) if qual == "*" else bytes(ord(c) - 33 for c in qual)
    return Rec(tid, pos0, mapq, 0, flag, len(seq), mtid, mpos0, isize, qname, cigar, seq4, q, aux, 0)
