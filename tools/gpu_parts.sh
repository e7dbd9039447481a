# border-only / interior-only diagnostic lines.  Usage: bash tools/gpu_parts.sh <tag> [configs...]
TAG=${1:-parts}; shift
CFGS=${@:-2 4 5}
mkdir -p gpurun_out
for c in $CFGS; do for p in 1 2; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --parts $p 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_config${c}_parts$p.json; done; done
