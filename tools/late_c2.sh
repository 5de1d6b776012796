#!/usr/bin/env bash
# BASELINE.json's C2 workload at full size through the CLI against tests/golden/c2/digests.json (MD5 + size of the reference's own
# outputs), without pytest: one line per output in gpurun_out/late_c2.log. Same check as
# tests/test_gpu_parity.py::test_c2_full_size_outputs_equal_the_reference_digests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/late_c2.log
: > $L
B=seeksv_b200/bin
T=$(mktemp -d)
want() { python3 - "$1" "$2" <<'PY'
import json, sys
d = json.load(open("tests/golden/c2/digests.json"))
for k in sys.argv[1].split("/"):
    d = d[k]
print(d[sys.argv[2]])
PY
}
chk() {  # name, file, digest path
  got=$(md5sum < "$2" | cut -d' ' -f1); sz=$(stat -c %s "$2")
  if [ "$got" = "$(want "$3" md5)" ] && [ "$sz" = "$(want "$3" bytes)" ]; then echo "ok   $1 ($sz bytes)" >> $L; else echo "DIFF $1 ($sz bytes, md5 $got)" >> $L; fi
}
SECONDS=0
$B/svsim --out $T/c2 --genome chr21:46709983 --cov 30 --nsv 500 --seed 20261017 2>/dev/null
echo "svsim ${SECONDS} s" >> $L
$B/seeksv getclip -o $T/o $T/c2.bam 2>/dev/null || echo "FAIL getclip rc=$?" >> $L
for e in clip.gz clip.fq.gz unmapped_1.fq.gz unmapped_2.fq.gz; do zcat $T/o.$e > $T/x; chk "getclip $e" $T/x ".$e"; done
$B/minialign $T/c2.fa $T/o.clip.fq.gz > $T/o.clip.sam
chk "clip.sam" $T/o.clip.sam "clip.sam"
$B/seeksv getsv $T/o.clip.sam $T/c2.bam $T/o.clip.gz $T/o.sv $T/o.unm > $T/o.out 2>/dev/null || echo "FAIL getsv rc=$?" >> $L
chk "getsv .sv" $T/o.sv "getsv/sv"; chk "getsv stdout" $T/o.out "getsv/stdout"
$B/seeksv somatic $T/c2.bam $T/o.clip.gz $T/o.sv $T/o.somatic 2>/dev/null || echo "FAIL somatic rc=$?" >> $L
chk "somatic (self)" $T/o.somatic "somatic (self)"
$B/seeksv getsv -n 0 -D $T/o.clip.sam $T/c2.bam $T/o.clip.gz $T/o.n.sv $T/o.unm > $T/o.n.out 2>/dev/null
chk "getsv -n 0 -D .sv" $T/o.n.sv "getsv -n 0 -D/sv"; chk "getsv -n 0 -D stdout" $T/o.n.out "getsv -n 0 -D/stdout"
echo "finished ${SECONDS} s" >> $L
cat $L
