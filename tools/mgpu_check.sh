#!/usr/bin/env bash
# N-rank check of seeksv_b200.mgpu on the fixtures: the four files of every sharding mode against the goldens.
# usage: tools/mgpu_check.sh N   (on a box with N GPUs)
set -euo pipefail
N=${1:-2}
cd "$(dirname "$0")/.."
out=$(mktemp -d)
CASES="fuzz/f11 fuzz/f12 micro/tumor example/cancer"; BYS="range chromosome"; SVCASES="fuzz/f11 micro/tumor example/cancer"; TRIOS="example:normal:cancer micro:normal:tumor"
if [ -n "${QUICK:-}" ]; then CASES="fuzz/f11"; BYS="range"; SVCASES="micro/tumor"; TRIOS="micro:normal:tumor"; fi   # (an N = 8 box is charged 8 x)
for case in $CASES; do
  for by in $BYS; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
      -m seeksv_b200.mgpu getclip --by $by -o "$out/o" tests/golden/$case.sort.bam 2>/dev/null
    for pair in clip.gz:clip.txt clip.fq.gz:clip.fq.txt unmapped_1.fq.gz:unmapped_1.fq.txt unmapped_2.fq.gz:unmapped_2.fq.txt; do
      cmp <(zcat "$out/o.${pair%%:*}") tests/golden/$case.${pair##*:} || { echo "MISMATCH $case $by $pair"; exit 1; }
    done
    echo "ok $case --by $by (N=$N)"
  done
done
# getsv and somatic on N ranks: additive passes on the shards' own records + collectives on device tensors
for case in $SVCASES; do
  zcat -f tests/golden/$case.clip.txt | gzip -1 > "$out/clip.gz"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
    -m seeksv_b200.mgpu getsv -- tests/golden/$case.clip.sam tests/golden/$case.sort.bam "$out/clip.gz" "$out/o.sv" "$out/o.unm" > "$out/o.stdout" 2>/dev/null
  cmp "$out/o.sv" tests/golden/$case.sv || { echo "MISMATCH getsv $case"; exit 1; }
  cmp "$out/o.stdout" tests/golden/$case.getsv.stdout || { echo "MISMATCH getsv stdout $case"; exit 1; }
  echo "ok getsv $case (N=$N)"
done
for trio in $TRIOS; do
  d=${trio%%:*}; rest=${trio#*:}; normal=${rest%%:*}; tumour=${rest##*:}
  zcat -f tests/golden/$d/$normal.clip.txt | gzip -1 > "$out/nclip.gz"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
    -m seeksv_b200.mgpu somatic -- tests/golden/$d/$normal.sort.bam "$out/nclip.gz" tests/golden/$d/$tumour.sv "$out/o.somatic" 2>/dev/null
  cmp "$out/o.somatic" tests/golden/$d/$tumour.somatic.temp.sv || { echo "MISMATCH somatic $d"; exit 1; }
  echo "ok somatic $d (N=$N)"
done
