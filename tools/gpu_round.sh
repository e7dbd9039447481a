# One GPU round: parity tests, bench lines (all configs), ncu launch list + one full capture of the dominant kernel
# (summarised on the box: the .ncu-rep with imported source exceeds what gpurun copies back).
# Usage: bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --config 2 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_config2.json
for c in 1 3 4 5 6 7 8 9; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_config$c.json; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_reference_config2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_config2.csv python bench.py --config 2 --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1
bash tools/gpu_ncu.sh ${TAG}_ncu_up2x_config2 resample_up2x 2 -- --config 2 --steps 1 --warmup 1 --no-cpu > /dev/null
./avisynth-jincresize_b200/fma_peak > gpurun_out/${TAG}_fma_peak.jsonl
