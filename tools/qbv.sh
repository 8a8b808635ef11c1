#!/bin/bash
# tools/qbv.sh <spp> <workload> <variant ...>: tools/qb.sh once per library variant under scratch/variants/ ("default" = in-tree)
spp=$1; wl=$2; shift; shift
for v in "$@"; do
  if [ "$v" = default ]; then lib=""; else lib="scratch/variants/$v/libpt_core.so"; fi
  PT_CORE_LIB=$lib tools/qb.sh $spp $wl $v
done
