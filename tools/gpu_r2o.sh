TAG=${1:-r02o}
mkdir -p gpurun_out
B="--no-cpu --no-all-configs --no-verify --plugin-threads 0 --bands 0"
run() { JINCRESIZE_B200_STRIP_PLAN=$3 timeout 600 python bench.py --config $1 --steps ${4:-20} --warmup 3 $B --parts $2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config',$1,'parts',$2,'plan',$3,'ms/step',round(d['ms_per_step'],4),'luma launch ms',round(r['launch_ms'],4),'frac',round(r['frac'],3))"; }
run 5 2 1 10; run 5 2 0 10
run 2 2 1; run 2 2 0
run 3 2 1; run 3 2 0
bash tools/gpu_ncu.sh ${TAG}_ncu_down_strips_config5 resample_down 2 -- --config 5 --steps 1 --warmup 1 $B --parts 2 > /dev/null
grep -E "gpu__time|pipe_fma_cycles|issue_active|launch__grid|bank_conflicts_pipe_lsu_mem_shared.sum|stall_" gpurun_out/${TAG}_ncu_down_strips_config5_summary.txt | cut -c1-130 | head -30
