# last checks of round 2 (third session) on the final tree: whole GPU suite, smoke, N=1 bench line, phase timings, launch list, ncu --set full
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2s3_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2s3_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2s3_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s3_bench_n1.json 2> gpurun_out/r2s3_bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2s3_bench_n1.json').read().strip().split(chr(10))[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'fused', d['e2e_fused']['ms_per_step'], 'step frac', d['roofline']['step']['frac'], 'walk frac', d['roofline']['frac'])
print(d['roofline']['kernels_ms_per_step'])
"
timeout 300 python tools/e2e_probe.py 4 > gpurun_out/r2s3_e2e_probe.log 2>&1
grep "ITER" gpurun_out/r2s3_e2e_probe.log | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2s3_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2s3_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'rec_walk|guess_starts|rs_sort|cluster_build|text_write|text_heads|rows_pass|unmapped_write|clip_eval|depth_marks|make_keys|cand_group|cj_' \
  --launch-skip 12 --launch-count 26 -o gpurun_out/r2s3_full -f \
  python bench.py --steps 1 --warmup 1 --value-only > gpurun_out/r2s3_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r2s3_full.ncu-rep
