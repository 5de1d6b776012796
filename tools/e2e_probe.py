"""In-process phase timings of the two commands on the C2 bench workload (run bench.py once first to generate it)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SEEKSV_B200_TIMING"] = "1"
import seeksv_b200 as S
W = os.environ.get("SEEKSV_B200_BENCH_DIR", "/tmp/seeksv_b200_bench")
bam = W + "/c2_chr21_46709983.bam"
sam = W + "/c2_chr21_46709983.clip.sam"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for it in range(n):
    t0 = time.perf_counter()
    S.run_cli(["getclip", "-o", W + "/probe", bam])
    t1 = time.perf_counter()
    S.run_cli(["getsv", sam, bam, W + "/probe.clip.gz", W + "/probe.sv", W + "/probe.unm"])
    t2 = time.perf_counter()
    print("ITER %d getclip %.1f ms getsv %.1f ms" % (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), file=sys.stderr)
