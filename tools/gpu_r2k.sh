# round 2, third session: text_heads split of the text writer, batched member chains in cluster_build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "candidate_order or getclip or c2_full or config3 or other_config or campaign or shards or run_keeps or long_clips" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2k_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().split(chr(10))[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'fused', d['e2e_fused']['ms_per_step'], 'step frac', d['roofline']['step']['frac'])
print(d['roofline']['kernels_ms_per_step'])
"
