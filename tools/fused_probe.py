"""Phase timings of `seeksv run -- getclip -- getsv` on the C2 bench workload (run bench.py once first)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SEEKSV_B200_TIMING"] = "1"
import seeksv_b200 as S
W = os.environ.get("SEEKSV_B200_BENCH_DIR", "/tmp/seeksv_b200_bench")
bam = W + "/c2_chr21_46709983.bam"
sam = W + "/c2_chr21_46709983.clip.sam"
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    t0 = time.perf_counter()
    S.run_cli(["run", "--", "getclip", "-o", W + "/probe", bam, "--", "getsv", sam, bam, W + "/probe.clip.gz", W + "/probe.sv", W + "/probe.unm"])
    print("ITER %d fused %.1f ms" % (it, 1e3 * (time.perf_counter() - t0)), file=sys.stderr)
