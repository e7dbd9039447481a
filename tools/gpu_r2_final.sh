# round 2, final code: GPU tests, the default bench line, the non-BASELINE workloads, the reference arm, the strip-plan and
# bulk-copy A/B lines.  Usage: bash tools/gpu_r2_final.sh <tag>
TAG=${1:-r02z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py 2> gpurun_out/${TAG}_bench_default.err | tail -1 > gpurun_out/${TAG}_bench_config2_default.json; cut -c1-300 gpurun_out/${TAG}_bench_config2_default.json
for c in 6 7 8 9; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --no-all-configs 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_config$c.json; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_reference_config2.json
B="--no-cpu --no-all-configs --plugin-threads 0 --bands 0"
# A/B lines (kernel-only): strips from the plan vs derived per block, float tiles through registers vs the bulk-copy engine
for c in 1 2 3 4 5 6 9; do for plan in 1 0; do JINCRESIZE_B200_STRIP_PLAN=$plan timeout 300 python bench.py --config $c --steps 20 --warmup 3 $B 2>/dev/null | tail -1 > gpurun_out/${TAG}_ab_config${c}_plan$plan.json; done; done
for tma in 0 1; do JINCRESIZE_B200_TMA=$tma timeout 300 python bench.py --config 4 --steps 20 --warmup 3 $B 2>/dev/null | tail -1 > gpurun_out/${TAG}_ab_config4_bulkcopy$tma.json; done
./avisynth-jincresize_b200/fma_peak > gpurun_out/${TAG}_fma_peak.jsonl
for f in gpurun_out/${TAG}_ab_*.json gpurun_out/${TAG}_bench_config[6-9].json; do python -c "
import json,sys
d=json.loads(open('$f').read()); r=d['roofline']; print('$f'.split('/')[-1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'launch ms', round(r['launch_ms'],4), 'frac', round(r['frac'],3), d.get('verified',{}).get('ok') if d.get('verified') else None)"; done
