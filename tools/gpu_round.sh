# One GPU round: parity tests, bench lines, ncu launch list + one full capture.  Usage: bash tools/gpu_round.sh <tag>
TAG=${1:-r01b}
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --config 2 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_config2.json
for c in 1 3 4 5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_config$c.json; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_config2.csv python bench.py --config 2 --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:resample_up2x -s 2 -c 1 -o gpurun_out/${TAG}_prof_up2x_c2 -f python bench.py --config 2 --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
