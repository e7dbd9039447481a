# round 2, call D: suite, default bench, the rational-ratio and general workloads
TAG=${1:-r02d}
mkdir -p gpurun_out
JINCRESIZE_B200_DEBUG=1 timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu_full.log 2>&1; tail -15 gpurun_out/${TAG}_pytest_gpu_full.log
(time timeout 900 python bench.py) 2> gpurun_out/${TAG}_bench_default.err | tail -1 > gpurun_out/${TAG}_bench_default.json
tail -4 gpurun_out/${TAG}_bench_default.err
for c in 6 7 8 9; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --bands 0 2>gpurun_out/${TAG}_bench_config$c.err | tail -1 > gpurun_out/${TAG}_bench_config$c.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_config$c.json"))
    print($c, round(d["value"]), round(d["roofline"]["frac"],3), d["roofline"]["kernel"][:30], "e2e", round(d["e2e"]["value"]), "pinned", round(d["e2e_pinned"]["value"]), d["verified"])
except Exception as e:
    print($c, "ERR", e); print(open("gpurun_out/${TAG}_bench_config$c.err").read()[-600:])
PY
done
