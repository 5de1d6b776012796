# last checks of the round on the final tree: whole GPU suite, smoke, N=1 bench line, phase timings of the two commands
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke_final.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench rc=$?"
timeout 300 python tools/e2e_probe.py 4 > gpurun_out/r2_e2e_probe_final.log 2>&1
grep "ITER\|load BAM\|read clip\|join\|device passes\|gzip" gpurun_out/r2_e2e_probe_final.log | tail -22
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n1_final.json').read().strip().split(chr(10))[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'fused', d['e2e_fused']['ms_per_step'], 'step frac', d['roofline']['step']['frac'], 'walk frac', d['roofline']['frac'])
"
