set -x
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/pytest_gpu_4.log
for c in 2 1 3 4; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu 2>&1 | tail -2 | tee gpurun_out/bench_c$c.json; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:resample_up2x -s 2 -c 1 -o gpurun_out/prof_up2x_c2 -f python bench.py --config 2 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
