#!/bin/bash
# One compact line of bench.py per (workload, env setting).  usage: tools/qb.sh <spp> <workload> [label]
spp=$1; wl=$2; label=${3:-$wl}
python bench.py --steps 1 --warmup 1 --spp $spp --workload $wl --no-cpu-baseline 2>gpurun_out/qb_err.log | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    k=d['roofline']['kernels']; p=d['per_ray']
    print('$label', 'Mrays/s=%.0f ms=%.1f' % (d['value'], d['ms_per_step']), ' '.join('%s=%.1fms' % (n, k[n]['ms_total']) for n in k), 'iters=%d' % k['extend']['launches'], 'nbox=%.1f/%.1f ntri=%.2f/%.2f' % (p['n_box_closest'], p['n_box_shadow'], p['n_tri_closest'], p['n_tri_shadow']), 'build=%.1fms nodes=%d' % (d['config']['bvh_build_ms'], d['config']['bvh_nodes']))
except Exception as e:
    print('$label', 'FAILED', e)
"
