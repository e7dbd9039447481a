set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"resample|axis|phase" -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --config 2 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"resample|axis|phase" -c 100 --csv --log-file gpurun_out/launches_c4.csv python bench.py --config 4 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launches4.log 2>&1
timeout 600 python bench.py --config 5 --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_c5.json
