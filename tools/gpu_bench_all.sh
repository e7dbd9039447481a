# Parity tests + kernel-only bench of every config.  Usage: bash tools/gpu_bench_all.sh <tag> [configs...]
TAG=${1:-all}; shift
CFGS=${@:-1 2 3 4 5}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --tb=short -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_gpu.log
for c in $CFGS; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_config$c.json
done
