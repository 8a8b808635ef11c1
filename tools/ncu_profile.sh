#!/bin/bash
# ncu evidence of a workload (run on the GPU box): (1) launch list of a whole small render, aggregated per kernel;
# (2) --set full of one mid-render launch each of k_extend, k_shade, k_shadow (one wavefront pool), summarised.
#   tools/ncu_profile.sh <workload> <spp> <tag>     -> gpurun_out/<tag>_launch_shares.txt, <tag>_ncu_summary.txt, <tag>.ncu-rep
set -e
w=$1; spp=$2; tag=$3
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python tools/render_once.py --workload $w --spp $spp --warm 0 --pools 2 > /dev/null 2>&1
python tools/launch_shares.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_shares.txt
rm -f gpurun_out/${tag}_launches.csv
ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_shade|k_shadow" --launch-skip 12 --launch-count 3 \
    -f -o gpurun_out/${tag} python tools/render_once.py --workload $w --spp $spp --warm 0 --pools 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${tag}.ncu-rep > gpurun_out/${tag}_ncu_summary.txt
head -30 gpurun_out/${tag}_launch_shares.txt
