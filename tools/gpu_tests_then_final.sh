cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2e_pytest.log
bash tools/gpu_final.sh
