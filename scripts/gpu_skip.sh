#!/bin/bash
# timing experiment: marginal cost of the phases of the frames kernel (results are wrong when a phase is skipped)
for wl in c2 c1f; do
for m in 0 1 2 4 8 6 7 14 15; do
  if [ $wl = c1f ]; then EXTRA="--workload c1 --path frames"; else EXTRA="--workload c2"; fi
  ASR_B200_DBG_SKIP=$m python bench.py $EXTRA --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl skip=$m kernel_ms %.4f'%d['roofline']['kernel_ms'])"
done; done
