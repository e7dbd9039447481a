# round 2, call B: full GPU suite, default bench line (all legs), reference arm
TAG=${1:-r02b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/${TAG}_smi.txt
nproc | tee -a gpurun_out/${TAG}_smi.txt
JINCRESIZE_B200_DEBUG=1 timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider --durations=12 > gpurun_out/${TAG}_pytest_gpu_full.log 2>&1; tail -40 gpurun_out/${TAG}_pytest_gpu_full.log
(time timeout 900 python bench.py) 2> gpurun_out/${TAG}_bench_default.err | tail -1 | tee gpurun_out/${TAG}_bench_default.json
tail -5 gpurun_out/${TAG}_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_reference.json
