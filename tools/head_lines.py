"""How much of a BAM's uncompressed stream the record heads touch at a given fetch granularity (analysis tool, CPU only).

    python tools/head_lines.py <file.bam> [max_uncompressed_MB=260]

The full-pass walkers (csrc/bam_index.cu, getclip.cu, getsv.cu) read the 4-byte block_size of every record (walk_count) or its
head - fixed part + read name + CIGAR (clip_walk, decode_walk); DRAM delivers whole lines. C2: heads touch 27.7 / 38.6 / 60.3 % of
the stream at 32 / 64 / 128 bytes, the block_size words alone 11.9 / 22.8 / 44.5 % - ncu's 1.21 GB for walk_count is exactly the
128-byte figure (44.5 % of 2.71 GB), so 128 bytes is the granularity that counts on this part.
"""
import struct
import sys
import zlib

import numpy as np


def main():
    cap = (int(sys.argv[2]) if len(sys.argv) > 2 else 260) * 1000000
    raw = open(sys.argv[1], "rb").read(cap // 3 + (1 << 20))
    o, tot, parts = 0, 0, []
    while o + 18 < len(raw) and tot < cap:
        bsize = struct.unpack_from("<H", raw, o + 16)[0] + 1
        if o + bsize > len(raw):
            break
        d = zlib.decompress(raw[o + 18:o + bsize - 8], -15)
        parts.append(d)
        tot += len(d)
        o += bsize
    s = b"".join(parts)
    p = 8 + struct.unpack_from("<i", s, 4)[0]
    nref = struct.unpack_from("<i", s, p)[0]
    p += 4
    for _ in range(nref):
        p += 8 + struct.unpack_from("<i", s, p)[0]
    offs, heads = [], []
    while p + 36 <= len(s):
        bs = struct.unpack_from("<i", s, p)[0]
        if p + 4 + bs > len(s):
            break
        offs.append(p)
        heads.append(36 + s[p + 12] + 4 * struct.unpack_from("<H", s, p + 16)[0])
        p += 4 + bs
    offs, heads = np.array(offs, dtype=np.int64), np.array(heads, dtype=np.int64)
    n = len(offs)
    span = offs[-1] + heads[-1] - offs[0]
    print("%d records, mean size %.1f bytes, mean head %.1f bytes" % (n, (offs[-1] - offs[0]) / (n - 1), heads.mean()))
    for gran in (32, 64, 128):
        first, last = offs // gran, (offs + heads - 1) // gran
        width = int((last - first).max()) + 1
        lines = np.unique(np.concatenate([np.minimum(first + k, last) for k in range(width)]))
        words = np.unique(np.concatenate([first, (offs + 3) // gran]))
        print("%3d-byte lines: heads touch %.1f %% of the stream (%.2f lines per record), the block_size words alone %.1f %%" % (
            gran, 100.0 * len(lines) * gran / span, len(lines) / n, 100.0 * len(words) * gran / span))
    per = np.bincount(offs >> 14)[(offs[0] >> 14) + 1:-1]
    print("records per 16 KiB chunk: mean %.1f, min %d, max %d" % (per.mean(), per.min(), per.max()))


if __name__ == "__main__":
    main()
