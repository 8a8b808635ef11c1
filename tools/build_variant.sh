#!/bin/bash
# Builds a variant of libpt_core.so with extra nvcc flags into scratch/variants/<name>/ (A/B
# experiments on the GPU box: PT_CORE_LIB=scratch/variants/<name>/libpt_core.so python bench.py ...).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/scratch/variants/$name
mkdir -p "$out"
cd "$root/path-tracing_b200/csrc"
for f in pt_core bvh_build textures wavefront postprocess unit_kernels; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
       -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -Wno-deprecated-gpu-targets "$@" \
       -c $f.cu -o "$out/$f.o" &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libpt_core.so" "$out"/*.o -cudart static
rm -f "$out"/*.o
echo "$out/libpt_core.so"
