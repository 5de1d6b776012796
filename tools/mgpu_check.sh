#!/usr/bin/env bash
# N-rank check of seeksv_b200.mgpu on the fixtures: the four files of every sharding mode against the goldens.
# usage: tools/mgpu_check.sh N   (on a box with N GPUs)
set -euo pipefail
N=${1:-2}
cd "$(dirname "$0")/.."
out=$(mktemp -d)
for case in fuzz/f11 fuzz/f12 micro/tumor example/cancer; do
  for by in range chromosome; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
      -m seeksv_b200.mgpu getclip --by $by -o "$out/o" tests/golden/$case.sort.bam 2>/dev/null
    for pair in clip.gz:clip.txt clip.fq.gz:clip.fq.txt unmapped_1.fq.gz:unmapped_1.fq.txt unmapped_2.fq.gz:unmapped_2.fq.txt; do
      cmp <(zcat "$out/o.${pair%%:*}") tests/golden/$case.${pair##*:} || { echo "MISMATCH $case $by $pair"; exit 1; }
    done
    echo "ok $case --by $by (N=$N)"
  done
done
