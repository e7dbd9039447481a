# round 2: the in-process multi-GPU tests and the bench under torchrun.  Usage (on an N-GPU box): bash tools/gpu_r2_multi.sh <tag> <N>
TAG=${1:-r02m}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv | tee gpurun_out/${TAG}_smi.txt
nvidia-smi topo -m 2>&1 | head -20 | tee -a gpurun_out/${TAG}_smi.txt
nproc | tee -a gpurun_out/${TAG}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rs > gpurun_out/${TAG}_pytest_gpu_full.log 2>&1; tail -12 gpurun_out/${TAG}_pytest_gpu_full.log
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -v --tb=short -p no:cacheprovider -k "two_gpu" 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_two_gpus.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu 2>gpurun_out/${TAG}_bench_gpus1.err | tail -1 | tee gpurun_out/${TAG}_bench_gpus1.json
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/${TAG}_bench_gpus$n.err | tail -1 | tee gpurun_out/${TAG}_bench_gpus$n.json
  fi
  tail -3 gpurun_out/${TAG}_bench_gpus$n.err
done
for c in 6 8 9; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --bands 0 --plugin-threads 0 2>gpurun_out/${TAG}_bench_config$c.err | tail -1 > gpurun_out/${TAG}_bench_config$c.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_config$c.json')); print($c, round(d['value']), round(d['roofline']['frac'],3), d['verified']['ok'])" || tail -5 gpurun_out/${TAG}_bench_config$c.err
done
