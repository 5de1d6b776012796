"""Times the BGZF inflate kernel alone on the C2 bench BAM (run bench.py once first: it generates the workload)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seeksv_b200 as S
import seeksv_b200.lib
W = os.environ.get("SEEKSV_B200_BENCH_DIR", "/tmp/seeksv_b200_bench")
bam = W + "/c2_chr21_46709983.bam"
image = open(bam, "rb").read()
ctx = S.Context(0)
ctx.prof(True)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ctx.prof_reset()
    t = time.perf_counter()
    out = S.lib.inflate_bgzf(ctx, image)
    dt = time.perf_counter() - t
    p = ctx.prof_read()["inflate_bgzf"]
    print("inflate_bgzf %.2f ms per launch (%d launches), %.1f GB/s out, wall %.0f ms" % (p["ms"] / p["launches"], p["launches"], len(out) / (p["ms"] / p["launches"]) / 1e6, dt * 1e3))
