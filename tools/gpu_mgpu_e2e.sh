# N-rank parity of the mgpu commands on the fixtures, then the bench with the phase timings of the multi-GPU end-to-end commands
N=${1:-2}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
QUICK=${QUICK:-} timeout 900 bash tools/mgpu_check.sh $N > gpurun_out/r2_mgpu_check_n$N.log 2>&1; echo "mgpu_check rc=$?"
tail -4 gpurun_out/r2_mgpu_check_n$N.log
SEEKSV_B200_TIMING=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_e2e_n$N.json 2> gpurun_out/r2_bench_e2e_n$N.err; echo "bench rc=$?"
grep "\[time\] mgpu" gpurun_out/r2_bench_e2e_n$N.err | tail -12
grep -v "^\[time\]\|^'" gpurun_out/r2_bench_e2e_n$N.err | tail -6
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_e2e_n$N.json').read().strip().split('\n')[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
"
