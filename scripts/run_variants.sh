#!/bin/bash
# development aid: bench several prebuilt variants of the library (asr-using-robust-nn_b200/build/variants/lib_<tag>.so)
cd asr-using-robust-nn_b200 && cp libasr_b200.so build/variants/lib_base.so
for V in "$@" base; do
  cp build/variants/lib_$V.so libasr_b200.so
  cd ..
  for WL in c1 c2; do
    python bench.py --workload $WL --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$V $WL kernel_ms %.4f ms/step %.4f'%(d['roofline']['kernel_ms'], d['ms_per_step']))"
  done
  cd asr-using-robust-nn_b200
done
