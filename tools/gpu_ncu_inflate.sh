# ncu --set full of the inflate kernel alone on the C2 image
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python - <<'PY' > /dev/null 2>&1
import os, sys
sys.path.insert(0, os.getcwd())
import bench
bench.ensure_tools()
os.makedirs(bench.WORK, exist_ok=True)
bench.make_bam(bench.WORK + "/c2_chr21_%d" % bench.C2_LEN, "chr21", bench.C2_LEN, bench.SEED, 500)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'inflate_bgzf' --launch-count 1 -o gpurun_out/${1:-r2_inflate_spec} -f \
  python tools/inflate_bench.py 1 > gpurun_out/r2_ncu_inflate.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_ncu_inflate.log
