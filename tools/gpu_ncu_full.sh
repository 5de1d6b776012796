# ncu --set full of the step's main kernels (one bench step after one warm-up step), report to gpurun_out/
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'rec_walk|guess_starts|rs_sort|cluster_build|text_write|rows_pass|unmapped_write|clip_eval|depth_marks' \
  --launch-skip 14 --launch-count 24 -o gpurun_out/r2_full -f \
  python bench.py --steps 1 --warmup 1 --value-only > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_ncu_full.log
