# on an N-GPU box: N-rank parity of mgpu getclip/getsv/somatic on the fixtures, then the partitioned bench
N=${1:-2}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 bash tools/mgpu_check.sh $N > gpurun_out/r2_mgpu_check_n$N.log 2>&1; echo "mgpu_check rc=$?"
tail -5 gpurun_out/r2_mgpu_check_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_n$N.err
