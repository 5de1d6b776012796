"""Host wall-clock of every ABI call of the resident-stream step (bench.py's device_step) on the C2 workload."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seeksv_b200 as S
from seeksv_b200 import lib as SL
W = os.environ.get("SEEKSV_B200_BENCH_DIR", "/tmp/seeksv_b200_bench")
bam_path = W + "/c2_chr21_46709983.bam"
sam = W + "/c2_chr21_46709983.clip.sam"
ctx = S.Context(0)
resident = S.Bam.from_bgzf(ctx, open(bam_path, "rb").read())
names, lens = resident.ref_names, resident.ref_lens
dptr, nbytes, first = resident.device_stream()
clip_gz = W + "/probe.clip.gz"
if not os.path.exists(clip_gz):
    S.run_cli(["getclip", "-o", W + "/probe", bam_path])
juncs, wins = S.plan_getsv(sam, clip_gz, names, lens)
nj, nw = len(juncs), len(wins)
j_arr = (SL.Junction * max(nj, 1))(*[SL.Junction(ut, up, dt, dp, us.encode(), ds.encode(), b"") for ut, up, us, dt, dp, ds in juncs])
w_arr = (SL.Window * max(nw, 1))(*[SL.Window(*w) for w in wins])
n_pos = sum(w[2] - w[1] + 1 for w in wins)
cnt = (C.c_int32 * max(nj, 1))()
dep = (C.c_int32 * max(n_pos, 1))()
acc = {}
def timed(name, fn):
    t = time.perf_counter()
    r = fn()
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
    return r
N = 20
for it in range(N + 3):
    if it == 3:
        acc.clear()
        t_all = time.perf_counter()
    b = timed("1 from_device", lambda: S.Bam.from_device(ctx, dptr, nbytes, first, len(names)))
    timed("2 set_refs", lambda: b.set_refs(names, lens))
    timed("3 getclip", lambda: b.getclip_sizes())
    timed("4 close", lambda: b.close())
    b = timed("5 from_device", lambda: S.Bam.from_device(ctx, dptr, nbytes, first, len(names)))
    timed("6 set_refs", lambda: b.set_refs(names, lens))
    n, tot, mean, sq = timed("7 insert_stats(+decode)", lambda: b.insert_stats(20, 5000000))
    timed("8 discordant", lambda: b.discordant_support_raw(j_arr, nj, SL.PairParams(20, mean, 25, 4), cnt))
    timed("9 window_depth", lambda: b.window_depth_raw(w_arr, nw, 20, dep))
    timed("a close", lambda: b.close())
total = (time.perf_counter() - t_all) / N
for k in sorted(acc):
    print("%-26s %7.3f ms" % (k, 1e3 * acc[k] / N))
print("step %.3f ms" % (1e3 * total))
