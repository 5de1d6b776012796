# inflate kernels: parity tests, timing of the kernel alone on the C2 image (both kernels), memcheck of the parity test
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "inflate" > gpurun_out/r2_inflate_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_inflate_pytest.log
python - <<'PY' > gpurun_out/r2_inflate_bench.log 2>&1
import os, sys
sys.path.insert(0, os.getcwd())
import bench
bench.ensure_tools()
os.makedirs(bench.WORK, exist_ok=True)
print(bench.make_bam(bench.WORK + "/c2_chr21_%d" % bench.C2_LEN, "chr21", bench.C2_LEN, bench.SEED, 500))
PY
for cfg in "SEEKSV_B200_INFLATE_CARVEOUT=100"; do
  echo "== $cfg" >> gpurun_out/r2_inflate_bench.log
  env $cfg timeout 300 python tools/inflate_bench.py >> gpurun_out/r2_inflate_bench.log 2>&1
done
cat gpurun_out/r2_inflate_bench.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -x -q -m gpu -k "inflate_refuses or inflate_matches" > gpurun_out/r2_inflate_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_inflate_memcheck.log
bash tools/gpu_ncu_inflate.sh r2_inflate_spec_v7
