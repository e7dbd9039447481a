# round 2, call F: suite, default bench, workloads 6-9, DRAM traffic of the batched dominant launch (configs 3, 4), ncu of the cells kernel
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu_full.log 2>&1; tail -8 gpurun_out/${TAG}_pytest_gpu_full.log
(time timeout 900 python bench.py) 2> gpurun_out/${TAG}_bench_default.err | tail -1 > gpurun_out/${TAG}_bench_default.json
tail -4 gpurun_out/${TAG}_bench_default.err
for c in 6 7 8 9; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --bands 0 2>gpurun_out/${TAG}_bench_config$c.err | tail -1 > gpurun_out/${TAG}_bench_config$c.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_config$c.json')); print($c, round(d['value']), round(d['roofline']['frac'],3), d['roofline']['kernel'][:24], 'e2e', round(d['e2e']['value']), 'pinned', round(d['e2e_pinned']['value']), d['verified']['ok'])" || tail -5 gpurun_out/${TAG}_bench_config$c.err
done
B="--no-cpu --no-all-configs --no-verify --plugin-threads 0 --bands 0"
for c in 3 4; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:resample_up2x -s 1 -c 1 --csv --log-file gpurun_out/${TAG}_traffic_config$c.csv python bench.py --config $c --steps 1 --warmup 1 $B > /dev/null 2>&1
  tail -3 gpurun_out/${TAG}_traffic_config$c.csv | cut -c1-60,200-400
done
bash tools/gpu_ncu.sh ${TAG}_ncu_cells_config6 resample_cells 2 -- --config 6 --steps 1 --warmup 1 $B > /dev/null
grep -E "gpu__time|pipe_fma_cycles|issue_active|bank_conflicts_pipe_lsu_mem_shared.sum|wavefronts_mem_shared" gpurun_out/${TAG}_ncu_cells_config6_summary.txt | cut -c1-120
