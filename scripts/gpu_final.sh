#!/bin/bash
# One gpurun call on the final tree: GPU parity tests, smoke, a bench line per BASELINE workload (+ the reference arm),
# ncu launch list + one full capture of the dominant kernel of the default step.   usage: scripts/gpu_final.sh <tag>
TAG=${1:-fin}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -2 $O/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 300 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
for WL in c1 c2b c3 c5 ref_vdr ref_sr; do
  timeout 300 python bench.py --workload $WL > $O/bench_$WL.json 2> $O/bench_$WL.err; echo "$WL rc=$?"
done
timeout 300 python bench.py --workload c1 --batch 1024 > $O/bench_c1_b1024.json 2> $O/bench_c1_b1024.err; echo "c1 b1024 rc=$?"
timeout 300 python bench.py --workload c4 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?"
python - <<PY
import json
for n in ("c2","reference","c1","c1_b1024","c2b","c3","c4","c5","ref_vdr","ref_sr"):
    try:
        d=json.loads(open("$O/bench_%s.json"%n).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(n, "value %.4g ms/step %.4f e2e %.4g"%(d["value"], d.get("ms_per_step") or 0, (d.get("e2e") or {}).get("value",0) or 0), "kernel_ms", r.get("kernel_ms"), "frac", r.get("frac"), "share", r.get("kernel_share_of_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 $CMD > $O/plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv $CMD > $O/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 200 $CMD > $O/plain2.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tile512" -s 3 -c 1 -f -o $O/prof_c2_tile512 $CMD > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"
