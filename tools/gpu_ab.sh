# quick kernel-only lines.  Usage: bash tools/gpu_ab.sh "<configs>" [parts]
B="--no-cpu --no-all-configs --plugin-threads 0 --bands 0"
for c in $1; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 $B --parts ${2:-3} 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config',$c,'parts',${2:-3},'value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'launch ms',round(r['launch_ms'],4),'frac',round(r['frac'],3),'verified',(d.get('verified') or {}).get('ok'))"; done
