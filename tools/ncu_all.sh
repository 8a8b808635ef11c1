#!/bin/bash
# ncu instruction / DRAM counters of every bench workload (run on the GPU box; writes gpurun_out/r2_kernel_counters.json)
#   tools/ncu_all.sh "chess:8 dragon:8 street:2 atrium:2"
set -e
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
echo "{" > gpurun_out/r2_kernel_counters.json
first=1
for ws in $1; do
  w=${ws%%:*}; s=${ws##*:}
  ncu --metrics $M --clock-control none --csv --log-file gpurun_out/ncu_$w.csv python tools/render_once.py --workload $w --spp $s --warm 0 --out gpurun_out/run_$w.json > gpurun_out/ncu_$w.log 2>&1
  [ $first = 1 ] || echo "," >> gpurun_out/r2_kernel_counters.json
  first=0
  echo "\"$w\": $(python tools/ncu_counters.py gpurun_out/ncu_$w.csv gpurun_out/run_$w.json)" >> gpurun_out/r2_kernel_counters.json
  rm -f gpurun_out/ncu_$w.csv
done
echo "}" >> gpurun_out/r2_kernel_counters.json
python -c "import json; d=json.load(open('gpurun_out/r2_kernel_counters.json')); [print(w, k, {x: round(v,1) if isinstance(v,float) else v for x,v in e.items()}) for w in d for k,e in d[w]['kernels'].items()]"
