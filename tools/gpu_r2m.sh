B="--no-cpu --no-all-configs --no-verify --plugin-threads 0 --bands 0"
run() { JINCRESIZE_B200_STAGGER_NS=$3 timeout 600 python bench.py --config $1 --steps 20 --warmup 3 $B --parts $2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config',$1,'parts',$2,'stagger',$3,'ms/step',round(d['ms_per_step'],4),'luma launch ms',round(r['launch_ms'],4),'frac',round(r['frac'],3))"; }
for s in 0 500 1000 2000 3000 5000; do run 2 3 $s; done
for s in 0 1000 2500; do run 3 3 $s; done
for s in 0 2000 6000; do run 4 3 $s; done
for s in 0 1000 2000; do run 1 3 $s; done
