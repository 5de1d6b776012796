# round 2, third session: device clip_join (kernels vs the CPU run of the rules, CLI with the device join), host vs device join
# inside the getsv command, cluster_build occupancy variants
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_clip_join.py -x -q -m gpu > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2i_pytest.log
for v in 0 12 16; do
  SEEKSV_B200_CLUSTER_MINB=$v timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench_minb$v.json 2> gpurun_out/r2i_bench_minb$v.err; echo "bench minb=$v rc=$?"
  python -c "
import json
d=json.loads(open('gpurun_out/r2i_bench_minb$v.json').read().strip().split(chr(10))[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'cluster_build', d['roofline']['kernels_ms_per_step'].get('cluster_build'))
"
done
timeout 300 python tools/e2e_probe.py 4 > gpurun_out/r2i_probe_host_join.log 2>&1
grep "ITER\|join" gpurun_out/r2i_probe_host_join.log | tail -8
SEEKSV_B200_DEVICE_JOIN=1 SEEKSV_B200_PROFILE=1 timeout 300 python tools/e2e_probe.py 4 > gpurun_out/r2i_probe_device_join.log 2>&1
grep "ITER\|join" gpurun_out/r2i_probe_device_join.log | tail -16
