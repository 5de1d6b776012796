"""Recall of the planted events of an svsim BAM in a seeksv call file.

    python tools/recall.py <prefix>.truth.tsv <calls.sv.txt> [tolerance=10]

svsim (tools/svsim.cpp) writes one line per planted event: type, chr, position, chr, position (the two breakpoints). An event
counts as found when one call has its two ends within `tolerance` bp of the two breakpoints (in either order). C2 (bench.py's
workload, 500 events): 296 / 296 DEL, 115 / 115 INV, 89 / 89 moved segments are found by the reference, whose output the CUDA
path reproduces byte for byte (tests/golden/c2/digests.json).
"""
import collections
import sys


def main():
    truth = [l.split("\t") for l in open(sys.argv[1]).read().splitlines() if l]
    tol = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    calls = []
    for l in open(sys.argv[2]).read().splitlines():
        if l and l[0] != "@":
            f = l.split("\t")
            calls.append((f[0], int(f[1]), f[4], int(f[5])))
    found, total = collections.Counter(), collections.Counter()
    for t in truth:
        ca, a, cb, b = t[1], int(t[2]), t[3], int(t[4])
        total[t[0]] += 1
        found[t[0]] += any((c[0] == ca and c[2] == cb and abs(c[1] - a) <= tol and abs(c[3] - b) <= tol) or
                           (c[0] == cb and c[2] == ca and abs(c[1] - b) <= tol and abs(c[3] - a) <= tol) for c in calls)
    for k in sorted(total):
        print("%s\t%d / %d" % (k, found[k], total[k]))
    print("calls\t%d" % len(calls))


if __name__ == "__main__":
    main()
