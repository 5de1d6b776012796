#!/usr/bin/env bash
# getsv of the campaign fixtures with the reference's literal per-position depth walk (SEEKSV_B200_LITERAL_DEPTH_WALK=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/late_literal.log
: > $L
G=tests/golden/fuzz
T=$(mktemp -d)
for s in f106 e3; do
  gzip -c $G/$s.clip.txt > $T/$s.clip.gz
  SEEKSV_B200_LITERAL_DEPTH_WALK=1 seeksv_b200/bin/seeksv getsv $G/$s.clip.sam $G/$s.sort.bam $T/$s.clip.gz $T/$s.sv $T/$s.unm > $T/$s.out 2>/dev/null
  if cmp -s $T/$s.sv $G/$s.sv && cmp -s $T/$s.out $G/$s.getsv.stdout; then echo "ok   literal walk $s" >> $L; else echo "DIFF literal walk $s" >> $L; fi
done
cat $L
