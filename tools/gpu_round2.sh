cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 700 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest3.log
timeout 300 python bench.py --steps 10 --warmup 3 --value-only > gpurun_out/r2_bench_value.json 2> gpurun_out/r2_bench_value.err; echo "bench rc=$?"
