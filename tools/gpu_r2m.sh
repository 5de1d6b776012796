# third session, last seconds of the GPU budget: the getclip-text tests that the 41-test subset did not hold, on the final tree
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 70 python -m pytest tests -x -q -m gpu -k "synthetic_cli or device_gzip_images or mgpu_entry or reads_sam or alternative_full_pass" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2m_pytest.log
