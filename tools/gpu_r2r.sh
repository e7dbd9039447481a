TAG=${1:-r02r2}
mkdir -p gpurun_out
JINCRESIZE_B200_TMA=1 timeout 900 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu_tma.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu_tma.log
B="--no-cpu --no-all-configs --plugin-threads 0 --bands 0"
run() { JINCRESIZE_B200_TMA=$3 timeout 300 python bench.py --config $1 --steps ${4:-20} --warmup 3 $B --parts $2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config',$1,'parts',$2,'tma',$3,'ms/step',round(d['ms_per_step'],4),'luma launch ms',round(r['launch_ms'],4),'frac',round(r['frac'],3), 'verified', d.get('verified'))"; }
run 4 3 0; run 4 3 1; run 4 1 0 ; run 4 1 1
