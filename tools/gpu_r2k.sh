TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu_full.log 2>&1; tail -6 gpurun_out/${TAG}_pytest_gpu_full.log
B="--no-cpu --no-all-configs --no-verify --plugin-threads 0 --bands 0"
for c in ${CONFIGS:-5 7}; do for p in 1 3; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 $B --parts $p 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config',$c,'parts',$p,'ms/step',round(d['ms_per_step'],4),'luma launch ms',round(r['launch_ms'],4),'frac',round(r['frac'],3))"; done; done
