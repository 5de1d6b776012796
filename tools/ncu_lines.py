"""Per-source-line summary of an `ncu --page source --csv --print-source sass,cuda` export: instructions executed, stall samples and
average active threads per CUDA source line (SASS rows are attributed to the last source line seen).

    ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > src.csv ; python tools/ncu_lines.py src.csv [top=40]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], newline="")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = {}
cur = None
src_of = {}
for r in rows:
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        i_inst, i_thr, i_samp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        i_ls, i_ss, i_w, i_br = hdr.index("stall_long_sb"), hdr.index("stall_short_sb"), hdr.index("stall_wait"), hdr.index("stall_branch_resolving")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] not in ("-", ""):
        cur = int(r[0])
        src_of[cur] = r[1]
        continue
    if cur is None:
        continue
    a = agg.setdefault(cur, [0, 0, 0, 0, 0, 0, 0])
    f = lambda i: int(float(r[i])) if r[i] not in ("", "-") else 0
    a[0] += f(i_inst); a[1] += f(i_thr); a[2] += f(i_samp); a[3] += f(i_ls); a[4] += f(i_ss); a[5] += f(i_w); a[6] += f(i_br)
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[2] for a in agg.values()) or 1
print("total warp instructions %.3e, samples %d" % (tot_i, tot_s))
print("%5s %7s %7s %6s  %6s %6s %6s %6s  %s" % ("line", "inst%", "samp%", "thr", "longsb", "shortsb", "wait", "branch", "source"))
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%5d %7.2f %7.2f %6.1f  %6.2f %6.2f %6.2f %6.2f  %s" % (ln, 100.0 * a[0] / tot_i, 100.0 * a[2] / tot_s, a[1] / max(1, a[0]),
          100.0 * a[3] / tot_s, 100.0 * a[4] / tot_s, 100.0 * a[5] / tot_s, 100.0 * a[6] / tot_s, src_of.get(ln, "")[:90]))
