# round 2, third session: gzip members of 64 KiB, P.clip.gz read back through the device inflate kernel in getsv
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "gzip or gz_reader or getsv or clip_join or run_keeps or mgpu or somatic" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().split(chr(10))[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'fused', d['e2e_fused']['ms_per_step'], 'step frac', d['roofline']['step']['frac'])
"
timeout 300 python tools/e2e_probe.py 5 > gpurun_out/r2j_probe.log 2>&1
grep "ITER\|time\]" gpurun_out/r2j_probe.log | tail -26
SEEKSV_B200_GZ_READ=host timeout 300 python tools/e2e_probe.py 4 > gpurun_out/r2j_probe_hostread.log 2>&1
grep "ITER" gpurun_out/r2j_probe_hostread.log | tail -3
