#!/bin/bash
# usage: scripts/gpu_prof.sh <tag> <workload> [kernel regex]   -> launch list + one full capture
TAG=$1; WL=$2; KR=${3:-tile512}
O=gpurun_out
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $O/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
$CMD > $O/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$KR -s 4 -c 1 -f -o $O/prof_$TAG $CMD > $O/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
python - <<PY
import csv,collections
rows=list(csv.reader(open("$O/launches_$TAG.csv")))
i=next(k for k,r in enumerate(rows) if r and r[0]=="ID")
t=collections.defaultdict(list)
for r in rows[i+1:]:
    if len(r)>14: t[r[4].split("(")[0][:60]].append(float(r[14]))
for k,v in sorted(t.items(), key=lambda kv:-sum(kv[1])): print(f"{k:62s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us")
PY
