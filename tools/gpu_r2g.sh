TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu_full.log 2>&1; tail -6 gpurun_out/${TAG}_pytest_gpu_full.log
for c in 6 9 8 7; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --bands 0 2>gpurun_out/${TAG}_bench_config$c.err | tail -1 > gpurun_out/${TAG}_bench_config$c.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_config$c.json')); print($c, round(d['value']), round(d['roofline']['frac'],3), d['roofline']['kernel'][:24], 'e2e', round(d['e2e']['value']), 'pinned', round(d['e2e_pinned']['value']), d['verified']['ok'])" || tail -5 gpurun_out/${TAG}_bench_config$c.err
done
