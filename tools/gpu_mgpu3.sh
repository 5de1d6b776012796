# third session of round 2: the partitioned bench at N ranks on the final tree (quick mgpu parity check first)
N=${1:-2}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 bash tools/mgpu_check.sh $N quick > gpurun_out/r2s3_mgpu_check_n$N.log 2>&1; echo "mgpu_check rc=$?"
tail -3 gpurun_out/r2s3_mgpu_check_n$N.log
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2s3_bench_n$N.json 2> gpurun_out/r2s3_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r2s3_bench_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/r2s3_bench_n$N.json').read().strip().split(chr(10))[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print(d.get('host_phases_ms'), d['roofline'].get('kernels_ms_per_step'))
"
