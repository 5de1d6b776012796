#!/usr/bin/env bash
# The fixtures that were added without GPU access, through the CLI, as fast as possible (no pytest, no python): one line per
# check in gpurun_out/late_checks.log. Used when only seconds of GPU time are left; tests/test_gpu_parity.py holds the same checks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/late_checks.log
: > $L
S=seeksv_b200/bin/seeksv
G=tests/golden
T=$(mktemp -d)
chk() { if cmp -s "$2" "$3"; then echo "ok   $1" >> $L; else echo "DIFF $1" >> $L; fi; }
for ds in "fuzz f106" "fuzz e3" "long lq"; do
  set -- $ds
  $S getclip -o $T/$2 $G/$1/$2.sort.bam 2>/dev/null || echo "FAIL getclip $2 rc=$?" >> $L
  for p in clip.gz:clip.txt clip.fq.gz:clip.fq.txt unmapped_1.fq.gz:unmapped_1.fq.txt unmapped_2.fq.gz:unmapped_2.fq.txt; do
    zcat $T/$2.${p%%:*} > $T/x.txt 2>/dev/null; chk "getclip $2 ${p%%:*}" $T/x.txt $G/$1/$2.${p##*:}
  done
  $S getsv $G/$1/$2.clip.sam $G/$1/$2.sort.bam $T/$2.clip.gz $T/$2.sv $T/$2.unm > $T/$2.out 2>/dev/null || echo "FAIL getsv $2 rc=$?" >> $L
  chk "getsv $2 .sv" $T/$2.sv $G/$1/$2.sv; chk "getsv $2 stdout" $T/$2.out $G/$1/$2.getsv.stdout
  $S somatic $G/$1/$2.sort.bam $T/$2.clip.gz $G/$1/$2.sv $T/$2.somatic 2>/dev/null || echo "FAIL somatic $2 rc=$?" >> $L
  chk "somatic $2" $T/$2.somatic $G/$1/$2.somatic.temp.sv
done
for s in f11 f12; do
  gzip -c $G/fuzz/$s.clip.txt > $T/$s.clip.gz
  $S somatic $G/fuzz/$s.sort.bam $T/$s.clip.gz $G/fuzz/$s.sv $T/$s.somatic 2>/dev/null || echo "FAIL somatic $s rc=$?" >> $L
  chk "somatic $s" $T/$s.somatic $G/fuzz/$s.somatic.temp.sv
done
if [ -x oracle/_ref/bamtool ]; then
  oracle/_ref/bamtool bam2sam $G/micro/tumor.sort.bam $T/tumor.sam 2>/dev/null
  $S getclip -o $T/sam $T/tumor.sam 2>/dev/null || echo "FAIL getclip sam rc=$?" >> $L
  zcat $T/sam.clip.gz > $T/x.txt 2>/dev/null; chk "getclip on SAM text clip.gz" $T/x.txt $G/micro/tumor.clip.txt
fi
echo "finished" >> $L
cat $L
