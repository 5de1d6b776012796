# third session: host-written gzip members of 64 KiB read back on the device (what mgpu getsv meets)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu -k "gz or gzip or mgpu or getsv_cli" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2l_pytest.log
