TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu_full.log 2>&1; tail -6 gpurun_out/${TAG}_pytest_gpu_full.log
B="--no-cpu --no-all-configs --no-verify --plugin-threads 0 --bands 0"
for c in 1 2 3; do for p in 1 3; do timeout 600 python bench.py --config $c --steps 20 --warmup 3 $B --parts $p 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config',$c,'parts',$p,'ms/step',round(d['ms_per_step'],4),'luma launch ms',round(r['launch_ms'],4),'frac',round(r['frac'],3))"; done; done
bash tools/gpu_ncu.sh ${TAG}_ncu_up2x_config1 resample_up2x 2 -- --config 1 --steps 1 --warmup 1 $B > /dev/null
grep -E "gpu__time|pipe_fma_cycles|issue_active|launch__grid" gpurun_out/${TAG}_ncu_up2x_config1_summary.txt | cut -c1-120
