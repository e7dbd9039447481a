# One `ncu --set full` capture of a kernel, summarised ON the box (the .ncu-rep with imported source can exceed the
# 64 MiB that gpurun copies back).  Usage: bash tools/gpu_ncu.sh <tag> <kernel-regex> <skip> -- <bench.py args...>
TAG=$1; KRE=$2; SKIP=$3; shift 4
mkdir -p gpurun_out
REP=/tmp/${TAG}.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o ${REP%.ncu-rep} -f python bench.py "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
ncu -i $REP --page details --csv > gpurun_out/${TAG}_details.csv 2>/dev/null
python tools/ncu_summary.py $REP 60 > gpurun_out/${TAG}_summary.txt 2>&1
ncu -i $REP --page source --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='Address')
h=rows[hi]; ci={n:i for i,n in enumerate(h)}
keep=['Address','Source','# Samples','Instructions Executed']+[k for k in h if k.startswith('stall_')]
w=csv.writer(sys.stdout); w.writerow(keep)
for r in rows[hi+1:]:
    if len(r)>ci['# Samples'] and float(r[ci['# Samples']] or 0)>0: w.writerow([r[ci[k]] for k in keep])
" > gpurun_out/${TAG}_source_sampled.csv
SZ=$(stat -c %s $REP); if [ $SZ -lt 30000000 ]; then cp $REP gpurun_out/; fi
cat gpurun_out/${TAG}_summary.txt | head -70
