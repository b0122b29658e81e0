#!/bin/bash
# usage: scripts/ab.sh "<env assignments>" workload...   -> prints value / ms_per_step / kernel_ms per workload
envs="$1"; shift
for w in "$@"; do
  env $envs python bench.py --workload $w --no-cpu-baseline --no-e2e > /tmp/ab_$w.json 2>/tmp/ab_$w.err || { echo "$w FAILED"; tail -3 /tmp/ab_$w.err; continue; }
  python - "$w" "$envs" <<PY
import json,sys
w=sys.argv[1]
d=json.loads(open(f"/tmp/ab_{w}.json").read().strip().splitlines()[-1])
print(f"[{sys.argv[2]}] {w}: value={d['value']:.4e} ms/step={d['ms_per_step']:.4f} kernel_ms={d['roofline']['kernel_ms']:.4f} frac={d['roofline']['frac']:.3f}")
PY
done
