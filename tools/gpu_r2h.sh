TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu_full.log 2>&1; tail -6 gpurun_out/${TAG}_pytest_gpu_full.log
(time timeout 900 python bench.py --no-cpu --bands 0) 2> gpurun_out/${TAG}_bench_default.err | tail -1 > gpurun_out/${TAG}_bench_default.json
tail -3 gpurun_out/${TAG}_bench_default.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_default.json"))
print("config2", round(d["value"]), round(d["roofline"]["frac"],3), round(d["whole_frame"]["frac"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["frac"],3))
for k,v in d["all_configs"].items(): print(k, round(v["value"]), round(v["frac"],3), round(v["whole_step_frac"],3), "e2e", round(v["e2e"]))
PY
for c in 1 3; do for p in 2; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --no-all-configs --bands 0 --plugin-threads 0 --parts $p 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('config',$c,'parts',$p,'ms/step',round(d['ms_per_step'],4))"; done; done
