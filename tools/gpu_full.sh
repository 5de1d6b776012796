# whole GPU suite + the N=1 bench + phase timings of the two commands
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2b}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench_n1.err
timeout 300 python tools/e2e_probe.py 3 > gpurun_out/${T}_e2e_probe.log 2>&1; echo "probe rc=$?"
tail -60 gpurun_out/${T}_e2e_probe.log
