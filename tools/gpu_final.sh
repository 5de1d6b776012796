# final evidence of the round: N=1 bench line, launch list of the same command, ncu --set full of the two dominant kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_final.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/r2_launches_final.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rec_walk|rs_sort|cluster_build|unmapped_pair|text_write|guess_starts|rows_pass' \
  --launch-skip 20 --launch-count 16 -o gpurun_out/r2_full_step_final -f python bench.py --steps 1 --warmup 1 --value-only > gpurun_out/r2_ncu_step_final.log 2>&1; echo "ncu step rc=$?"
bash tools/gpu_ncu_inflate.sh r2_inflate_spec_final
