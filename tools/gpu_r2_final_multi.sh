# round 2, final code, on an N-GPU box: the two-GPU tests and the bench under torchrun.  Usage: bash tools/gpu_r2_final_multi.sh <tag> <N>
TAG=${1:-r02z}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -v --tb=short -p no:cacheprovider -k "two_gpu" 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest_two_gpus.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/${TAG}_bench_gpus$N.err | tail -1 > gpurun_out/${TAG}_bench_config2_gpus$N.json
tail -3 gpurun_out/${TAG}_bench_gpus$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_config2_gpus$N.json"))
print("N",$N,"value",round(d["value"]),"frac",round(d["roofline"]["frac"],3),"e2e",round(d["e2e"]["value"]),"frac",round(d["e2e"]["frac"],3),"ceil GB/s",round(d["e2e"]["ceiling_gbs"],1),"pinned",round(d["e2e_pinned"]["value"]))
print("  bands",d["row_bands"])
PY
