# round 2, call A: full GPU suite (new pipeline tests, full-size parity), stale-registration test in its own process, bench lines
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/${TAG}_smi.txt
nproc | tee -a gpurun_out/${TAG}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider --deselect tests/test_gpu_pipeline.py::test_stale_registration_is_detected --durations=15 2>&1 | tail -80 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=short -p no:cacheprovider -k stale 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest_stale.log
for c in 2 3; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_config$c.json
done
JINCRESIZE_B200_HOSTREG=0 timeout 600 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_config2_nohostreg.json
