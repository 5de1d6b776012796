"""Token statistics of the DEFLATE streams inside a BGZF file - what the device inflate kernel (csrc/inflate.cu) spends its time on.

    python tools/deflate_stats.py <file.bam> [n_blocks=60] [seed=1]

A pure-Python inflate of a random sample of BGZF blocks (checked against zlib), counting per token: literal or match, the length
of its Huffman code(s) - codes longer than the kernel's single-lookup tables (10 bits literal/length, 8 bits distance) take the
canonical slow path -, extra bits, match lengths around the kernel's per-lane copy limit (16) and how often a match depends on
the output of the same 32-token batch. Analysis tool only (CPU, no GPU); numbers for the C2 workload are in profiles/r1_summary.md.
"""
import collections
import random
import struct
import sys
import zlib

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193,
             12289, 16385, 24577]
DIST_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]
CL_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


class Bits:
    def __init__(self, data):
        self.d, self.p = data, 0

    def get(self, n):
        v = 0
        for i in range(n):
            v |= ((self.d[self.p >> 3] >> (self.p & 7)) & 1) << i
            self.p += 1
        return v


def make_decoder(lens):
    """canonical Huffman (RFC 1951 3.2.2): {(length, code): symbol}"""
    count = collections.Counter(l for l in lens if l)
    code, nxt = 0, {}
    for bits in range(1, 16):
        code = (code + count.get(bits - 1, 0)) << 1
        nxt[bits] = code
    table = {}
    for s, l in enumerate(lens):
        if l:
            table[(l, nxt[l])] = s
            nxt[l] += 1
    return table


def decode_sym(b, table):
    code = 0
    for l in range(1, 16):
        code = (code << 1) | b.get(1)
        s = table.get((l, code))
        if s is not None:
            return s, l
    raise ValueError("bad code")


def inflate_stats(data, st):
    b, out = Bits(data), bytearray()
    n_blocks = 0
    while True:
        final, typ = b.get(1), b.get(2)
        n_blocks += 1
        st["deflate_blocks_type%d" % typ] += 1
        if typ == 0:
            b.p = (b.p + 7) & ~7
            ln = b.get(16)
            b.get(16)
            out += data[b.p >> 3:(b.p >> 3) + ln]
            b.p += 8 * ln
        else:
            if typ == 1:
                lit_lens = [8] * 144 + [9] * 112 + [7] * 24 + [8] * 8
                dist_lens = [5] * 30
            else:
                hlit, hdist, hclen = b.get(5) + 257, b.get(5) + 1, b.get(4) + 4
                cl = [0] * 19
                for i in range(hclen):
                    cl[CL_ORDER[i]] = b.get(3)
                clt = make_decoder(cl)
                lens = []
                while len(lens) < hlit + hdist:
                    s, _ = decode_sym(b, clt)
                    if s < 16:
                        lens.append(s)
                    elif s == 16:
                        lens += [lens[-1]] * (3 + b.get(2))
                    elif s == 17:
                        lens += [0] * (3 + b.get(3))
                    else:
                        lens += [0] * (11 + b.get(7))
                lit_lens, dist_lens = lens[:hlit], lens[hlit:]
                st["header_bits"] += b.p     # (position after the first header; one block per member is the rule)
            lt, dt = make_decoder(lit_lens), make_decoder(dist_lens)
            batch_start, in_batch = len(out), 0
            prev_match = False
            while True:
                s, l = decode_sym(b, lt)
                if s == 256:
                    break
                if in_batch == 32:
                    batch_start, in_batch = len(out), 0
                in_batch += 1
                st["tokens"] += 1
                st["litlen_code_len_%02d" % l] += 1
                if l > 10:
                    st["litlen_slow_path"] += 1
                if s < 256:
                    out.append(s)
                    st["literals"] += 1
                    prev_match = False
                    continue
                i = s - 257
                ln = LEN_BASE[i] + b.get(LEN_EXTRA[i])
                ds, dl = decode_sym(b, dt)
                dist = DIST_BASE[ds] + b.get(DIST_EXTRA[ds])
                st["matches"] += 1
                st["match_bytes"] += ln
                st["dist_code_len_%02d" % dl] += 1
                if dl > 8:
                    st["dist_slow_path"] += 1
                st["token_bits_sum"] += l + LEN_EXTRA[i] + dl + DIST_EXTRA[ds]
                st["match_len_le16" if ln <= 16 else "match_len_gt16"] += 1
                if dist < len(out) - batch_start + ln:          # source reaches into this batch's output (or overlaps itself)
                    st["match_batch_dependent"] += 1
                if prev_match:
                    st["match_after_match"] += 1
                prev_match = True
                for _ in range(ln):
                    out.append(out[-dist])
        if final:
            break
    st["deflate_blocks"] += n_blocks
    return bytes(out)


def main():
    path = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    rng = random.Random(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
    raw = open(path, "rb").read()
    offs, o = [], 0
    while o < len(raw):
        bsize = struct.unpack_from("<H", raw, o + 16)[0] + 1
        offs.append((o, bsize))
        o += bsize
    st = collections.Counter()
    picks = rng.sample(offs[1:-1], min(n, len(offs) - 2))
    for o, bsize in picks:
        payload = raw[o + 18:o + bsize - 8]
        got = inflate_stats(payload, st)
        assert got == zlib.decompress(payload, -15)
        st["bgzf_blocks"] += 1
        st["out_bytes"] += len(got)
        st["in_bytes"] += len(payload)
    t = st["tokens"]
    print("%d BGZF blocks of %d sampled: %d -> %d bytes, %d deflate blocks (%s)" % (
        st["bgzf_blocks"], len(offs), st["in_bytes"], st["out_bytes"], st["deflate_blocks"],
        ", ".join("type %d: %d" % (k, st["deflate_blocks_type%d" % k]) for k in range(3))))
    print("tokens %d (%.0f per block): %.1f %% matches, mean match length %.1f, %.1f %% of the output bytes from matches" % (
        t, t / st["bgzf_blocks"], 100.0 * st["matches"] / t, st["match_bytes"] / max(1, st["matches"]),
        100.0 * st["match_bytes"] / st["out_bytes"]))
    print("matches: %.1f %% of length <= 16, %.1f %% depend on the output of their own 32-token batch, %.1f %% follow a match" % (
        100.0 * st["match_len_le16"] / max(1, st["matches"]), 100.0 * st["match_batch_dependent"] / max(1, st["matches"]),
        100.0 * st["match_after_match"] / max(1, st["matches"])))
    print("literal/length codes longer than 10 bits: %.2f %% of the tokens; distance codes longer than 8 bits: %.2f %% of the matches" % (
        100.0 * st["litlen_slow_path"] / t, 100.0 * st["dist_slow_path"] / max(1, st["matches"])))
    for name in ("litlen_code_len", "dist_code_len"):
        tot = sum(v for k, v in st.items() if k.startswith(name))
        print(name + ": " + "  ".join("%d:%.1f%%" % (int(k[-2:]), 100.0 * v / tot) for k, v in sorted(st.items()) if k.startswith(name)))
    print("bits per match token %.1f; dynamic header %.0f bits per block = %.1f %% of the input" % (
        st["token_bits_sum"] / max(1, st["matches"]), st["header_bits"] / max(1, st["bgzf_blocks"]),
        100.0 * st["header_bits"] / (8.0 * st["in_bytes"])))


if __name__ == "__main__":
    main()
