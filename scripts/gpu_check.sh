#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench lines, ncu launch list + one full capture of the MFCC kernel.
# usage: scripts/gpu_check.sh <tag> [skip_ncu]
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log
tail -3 $O/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_$TAG.log
python bench.py --workload c1 --no-cpu-baseline > $O/bench_c1_$TAG.json 2> $O/bench_c1_$TAG.err; echo "c1 rc=$?"
python bench.py > $O/bench_c2_$TAG.json 2> $O/bench_c2_$TAG.err; echo "c2 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"
python - <<PY
import json
for n in ("c1","c2","ref"):
    try:
        d=json.loads(open("$O/bench_%s_$TAG.json"%n).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(n, "value %.3g e2e %.3g"%(d["value"], (d.get("e2e") or {}).get("value",0)), "kernel_ms", r.get("kernel_ms"), "fp32 frac", (r.get("fp32") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
if [ -z "$2" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $O/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
$CMD > $O/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tile512|frames512|mfcc_kernel" -s 3 -c 1 -f -o $O/prof_$TAG $CMD > $O/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
fi
