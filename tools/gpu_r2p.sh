TAG=${1:-r02p2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu_full.log 2>&1; tail -6 gpurun_out/${TAG}_pytest_gpu_full.log
B="--no-cpu --no-all-configs --no-verify --plugin-threads 0 --bands 0"
run() { JINCRESIZE_B200_STRIP_PLAN=$3 timeout 600 python bench.py --config $1 --steps ${4:-20} --warmup 3 $B --parts $2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config',$1,'parts',$2,'plan',$3,'ms/step',round(d['ms_per_step'],4),'luma launch ms',round(r['launch_ms'],4),'frac',round(r['frac'],3))"; }
run 5 2 1 10; run 5 2 0 10; run 5 3 1 10; run 5 3 0 10
run 4 2 1; run 4 2 0; run 4 3 1
run 9 2 1; run 9 2 0; run 9 3 1
for c in 1 2 3; do run $c 3 1; done
