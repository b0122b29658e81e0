#!/bin/bash
# usage: scripts/gpu_quick.sh <tag>   -> tiles check, c1/c2 bench lines (no CPU baseline), both tile shapes
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 300 python scripts/tiles_check.py > $O/tiles_check_$TAG.log 2>&1; echo "check rc=$?"; grep -c "status_eq=True" $O/tiles_check_$TAG.log; awk '{print $4}' $O/tiles_check_$TAG.log | sort -u | tail -3
for NW in 8 16; do
for WL in c1 c2; do
ASR_B200_TILE_WARPS=$NW timeout 300 python bench.py --workload $WL --no-cpu-baseline --no-e2e > $O/bench_${WL}_nw${NW}_$TAG.json 2> $O/bench_${WL}_nw${NW}_$TAG.err; echo "$WL nw$NW rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_${WL}_nw${NW}_$TAG.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("$WL nw$NW value %.4g ms/step %.4f kernel_ms %.4f hbm frac %.3f fp32 frac %.3f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["fp32"]["frac"]))
PY
done
done
