# ncu --set full of the candidate-side kernels of one bench step (after one warm-up step)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'cluster_build|text_write|text_heads|clip_eval|make_keys|cand_group' \
  --launch-skip 6 --launch-count 6 -o gpurun_out/r2_cand -f \
  python bench.py --steps 1 --warmup 1 --value-only > gpurun_out/r2_ncu_cand.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_ncu_cand.log
ls -la gpurun_out/r2_cand.ncu-rep
