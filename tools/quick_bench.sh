#!/bin/bash
# Runs bench.py (short) once per library variant and prints one compact line each.
# usage: tools/quick_bench.sh <spp> <variant|default> ...   (env vars pass through)
spp=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then lib=""; else lib="scratch/variants/$v/libpt_core.so"; fi
  PT_CORE_LIB=$lib python bench.py --steps 1 --warmup 1 --spp $spp --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
v='$v'
try:
    d=json.loads(sys.stdin.read())
    k=d['roofline']['kernels']
    print(v, 'Mrays/s=%.0f ms=%.1f' % (d['value'], d['ms_per_step']), ' '.join('%s=%.1fms' % (n, k[n]['ms_total']) for n in k), 'iters=%d' % k['extend']['launches'])
except Exception as e:
    print(v, 'FAILED', e)
"
done
