# round 2: the bench under torchrun at N GPUs of one box (what the driver's scaling run does).  Usage: bash tools/gpu_r2_scale.sh <tag> <N...>
TAG=${1:-r02s}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader | head -8
nproc; free -g | head -2
for n in "$@"; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --steps 20 --warmup 3 ) 2>gpurun_out/${TAG}_bench_gpus$n.err | tail -1 > gpurun_out/${TAG}_bench_gpus$n.json
  tail -4 gpurun_out/${TAG}_bench_gpus$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_gpus$n.json"))
    print("N",$n,"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["e2e"]["frac"],3),"ceil GB/s",round(d["e2e"]["ceiling_gbs"],1),"pinned",round(d["e2e_pinned"]["value"]),round(d["e2e_pinned"]["frac"],3))
    print("  bands",d["row_bands"])
    for k,v in (d["all_configs"] or {}).items():
        print("  ",k, round(v["value"]), round(v["frac"],3), "e2e", round(v["e2e"]) if v["e2e"] else None)
except Exception as e:
    print("ERR",e)
PY
done
