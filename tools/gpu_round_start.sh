#!/usr/bin/env bash
# First GPU call of a round (one B200):  gpurun --timeout 1500 -- 'bash tools/gpu_round_start.sh'
# Everything the CPU-only sessions could not run, in the order that matters: the parity suite (the tests added without a GPU are the
# last ones in tests/test_gpu_parity.py), the fuzz campaign through the CUDA path, one bench line, the launch list.
# Outputs land in gpurun_out/ (copy what is to be judged into profiles/).
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/gpu_tests.log
# reference vs oracle vs CUDA path on fresh fuzz seeds (oracle/_ref travels with the snapshot; bwa does not: minialign)
timeout 600 python tools/fuzz_campaign.py --gpu --seeds 1000:1040 --options 2 2>&1 | grep -v ": ok" | tee gpurun_out/fuzz_gpu.log
timeout 400 python tools/fuzz_campaign.py --gpu --edge --seeds 1040:1070 --connect 160 2>&1 | grep -v ": ok" | tee -a gpurun_out/fuzz_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# ceilings for the walkers: streaming read vs scattered / dependent 128-byte line reads
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/scatter_bw tools/scatter_bw.cu && /tmp/scatter_bw 2.7 | tee gpurun_out/scatter_bw.jsonl
echo done
