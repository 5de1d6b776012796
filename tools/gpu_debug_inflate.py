"""Debug aid: device inflate of BGZF files vs zlib; reports the mismatching regions per BGZF block."""
import gzip, os, struct, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seeksv_b200 as S
from seeksv_b200.lib import inflate_bgzf
ctx = S.Context(0)
for path in sys.argv[1:]:
    img = open(path, "rb").read()
    want = gzip.decompress(img)
    try:
        got = inflate_bgzf(ctx, img)
    except Exception as e:
        print(path, "ERROR", e)
        continue
    if got == want:
        print(path, "ok", len(want))
        continue
    # block table
    o, u, blocks = 0, 0, []
    while o < len(img):
        bs = struct.unpack_from("<H", img, o + 16)[0] + 1
        isize = struct.unpack_from("<I", img, o + bs - 4)[0]
        blocks.append((u, isize))
        u += isize
        o += bs
    print(path, "DIFFERS", len(got), len(want))
    for bi, (u0, n) in enumerate(blocks):
        g, w = got[u0:u0 + n], want[u0:u0 + n]
        if g == w:
            continue
        bad = [i for i in range(n) if g[i] != w[i]]
        runs, start, prev = [], bad[0], bad[0]
        for i in bad[1:]:
            if i != prev + 1:
                runs.append((start, prev + 1)); start = i
            prev = i
        runs.append((start, prev + 1))
        print(" block", bi, "ulen", n, "mismatching bytes", len(bad), "runs", len(runs), "first runs", runs[:8])
        a = runs[0][0]
        print("   got ", g[max(0, a - 8):a + 24].hex())
        print("   want", w[max(0, a - 8):a + 24].hex())
