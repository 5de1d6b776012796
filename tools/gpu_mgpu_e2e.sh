# N-rank bench with the phase timings of the multi-GPU end-to-end commands
N=${1:-2}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SEEKSV_B200_TIMING=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_e2e_n$N.json 2> gpurun_out/r2_bench_e2e_n$N.err; echo "bench rc=$?"
grep "\[time\]" gpurun_out/r2_bench_e2e_n$N.err | tail -40
tail -2 gpurun_out/r2_bench_e2e_n$N.err
