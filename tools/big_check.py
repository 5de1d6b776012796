"""A stream larger than 4 GiB through the whole-file path and through chromosome / range shards (each below 4 GiB): the
outputs must agree byte for byte - a check of the 64-bit offsets end to end. Usage: big_check.py [genome spec]"""
import gzip, hashlib, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import seeksv_b200 as S
from seeksv_b200 import sharding
W = "/tmp/seeksv_b200_big"
os.makedirs(W, exist_ok=True)
spec = sys.argv[1] if len(sys.argv) > 1 else "chr1:80000000,chr2:70000000,chr3:60000000"
pre = W + "/big"
if not os.path.exists(pre + ".bam"):
    t = time.perf_counter()
    subprocess.run([os.path.join(ROOT, "seeksv_b200", "bin", "svsim"), "--out", pre, "--genome", spec, "--cov", "30", "--nsv", "900",
                    "--seed", "7"], check=True, stderr=subprocess.DEVNULL)
    print("svsim %.0f s, BAM %.2f GB" % (time.perf_counter() - t, os.path.getsize(pre + ".bam") / 1e9), flush=True)
t = time.perf_counter()
assert S.run_cli(["getclip", "-o", W + "/whole", pre + ".bam"]) == 0
t_whole = time.perf_counter() - t
ctx = S.Context(0)
probe = S.Bam.open_refs(ctx, pre + ".bam", 0, 0)
names, lens = probe.ref_names, probe.ref_lens
probe.close()
def md5_gz(p):
    h = hashlib.md5()
    with gzip.open(p, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()
want = [md5_gz(W + "/whole" + e) for e in (".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz")]
def md5s(texts):
    return [hashlib.md5(t.encode("latin-1")).hexdigest() for t in texts]
# chromosome shards, one after the other
world = len(names)
parts, lasts, n_rec, nbytes = [], [], 0, 0
for r in range(world):
    w = sharding.open_ref_shard(ctx, pre + ".bam", r, world)
    lasts.append(w.last_mapped_tid())
    w.close()
prev = sharding.prev_tids(lasts)
for r in range(world):
    w = sharding.open_ref_shard(ctx, pre + ".bam", r, world)
    n_rec += w.bam.n_records
    nbytes += w.bam.device_stream()[1]
    parts.append(w.getclip(prev[r]))
    if r == world - 1:
        u1, u2 = w.pair_unmapped(b"".join(p[4] for p in parts))
    w.close()
got = md5s(["".join(p[0] for p in parts), "".join(p[1] for p in parts), u1, u2])
print("records %d, uncompressed %.2f GB, whole-file getclip %.2f s (%.1f M records/s)" % (n_rec, nbytes / 1e9, t_whole, n_rec / t_whole / 1e6))
print("chromosome shards == whole file:", got == want, flush=True)
# range shards (5 ranks), one after the other
plans = sharding.plan_range_shards(pre + ".bam", None, len(names), 5)
parts = []
for p in plans:
    w = sharding.RangeShardWorker(ctx, pre + ".bam", p)
    assert w.context_has_mapped_record()
    parts.append(w.getclip())
    if p is plans[-1]:
        u1, u2 = w.pair_unmapped(b"".join(x[4] for x in parts))
    w.close()
clip, fq = sharding.merge_range_texts([(x[0], x[1]) for x in parts])
got = md5s([clip, fq, u1, u2])
print("range shards == whole file:", got == want)
ctx.close()
assert got == want
