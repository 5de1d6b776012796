cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest2.log
timeout 300 python bench.py --steps 10 --warmup 3 --value-only > gpurun_out/r2_bench_value.json 2> gpurun_out/r2_bench_value.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --value-only > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu rc=$?"
