#!/usr/bin/env bash
# Build the UNMODIFIED reference seeksv (v1.2.3) from the sources where they lie under
# /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).
#
# Test infrastructure only: the product never executes anything under oracle/.
#
# Why a scratch copy is needed:
#   * five non-void functions fall off their end (clip_reads.h:79,83,219; clip_reads.cpp:558,570);
#     g++ >= 8 turns that into __builtin_unreachable and the binary traps at run time. The fix is a
#     `return 0;` in each — applied by sed to a throw-away copy under $TMPDIR, never to the repo.
#   * sam/libbam.a is non-PIC, hence -no-pie.
# Nothing but the linked binaries is written below oracle/_ref/.
set -euo pipefail
REF=${SEEKSV_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF/seeksv" ]; then
  echo "[build_ref] $REF not present (GPU box?) - keeping prebuilt oracle/_ref as is" >&2
  exit 0
fi
mkdir -p "$OUT"
if [ -x "$OUT/seeksv" ] && [ "$OUT/seeksv" -nt "$HERE/build_ref.sh" ] && [ -x "$OUT/bamtool" ] && [ "$OUT/bamtool" -nt "$HERE/bamtool.c" ] \
   && [ -x "$OUT/svcompare" ] && [ "$OUT/svcompare" -nt "$HERE/build_ref.sh" ]; then
  exit 0
fi
# the stand-alone evaluator (one source file + junction.h; its Makefile is `g++ svcompare.cpp -o svcompare`)
(cd "$REF/svcompare" && g++ -O2 -w svcompare.cpp -o "$OUT/svcompare")
T=$(mktemp -d)
trap 'rm -rf "$T"' EXIT
mkdir "$T/seeksv"
ln -s "$REF/sam" "$T/sam"
for f in "$REF"/seeksv/*.cpp "$REF"/seeksv/*.h "$REF"/seeksv/*.C; do
  [ "$(basename "$f")" = main.cpp ] && continue
  cat "$f" > "$T/seeksv/$(basename "$f")"
done
cd "$T/seeksv"
# guard: make sure we patch the lines we think we patch
sed -n 79p clip_reads.h | grep -q 'bool set_used(int u) { used = u; }'
sed -n 83p clip_reads.h | grep -q 'bool support_read_no_increase() { ++support_read_no; }'
sed -n 219p clip_reads.h | grep -q '^}'
sed -n 558p clip_reads.cpp | grep -q '^}'
sed -n 570p clip_reads.cpp | grep -q '^}'
sed -i -e '79s/used = u; }/used = u; return 0; }/' \
       -e '83s/++support_read_no; }/++support_read_no; return 0; }/' \
       -e '219s/^}/\treturn 0;\n}/' clip_reads.h
sed -i -e '570s/^}/\treturn 0;\n}/' -e '558s/^}/\treturn 0;\n}/' clip_reads.cpp
g++ -O2 -w -no-pie bam2depth.cpp cluster.cpp gzstream.C seeksv.cpp clip_reads.cpp getsv.cpp somatic.cpp \
    process_bwasw.cpp -o "$OUT/seeksv" -lz -lpthread -lm -L../sam -lbam
# helper linked against the same libbam: sam->bam conversion and .bai building (the bundled samtools
# binary needs libncurses.so.5 and does not run in this image)
g++ -O2 -w -no-pie -I"$REF/sam" "$HERE/bamtool.c" -x none -o "$OUT/bamtool" -L"$REF/sam" -lbam -lz -lpthread -lm
echo "[build_ref] built $OUT/seeksv, $OUT/bamtool and $OUT/svcompare" >&2
