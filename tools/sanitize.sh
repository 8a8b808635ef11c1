#!/bin/bash
# compute-sanitizer passes over the GPU tests (run on the GPU box: gpurun -- 'tools/sanitize.sh memcheck').
# Usage: tools/sanitize.sh memcheck|initcheck|racecheck [pytest args...]; logs to gpurun_out/sanitize_<tool>.log
set -u
tool=${1:-memcheck}
shift || true
tests=${*:-tests/test_gpu_units.py tests/test_gpu_textures.py tests/test_gpu_postprocess.py tests/test_gpu_scene_update.py
        tests/test_gpu_skinning.py tests/test_gpu_bc.py tests/test_gpu_traversal.py tests/test_gpu_render.py tests/test_gpu_configs.py}
mkdir -p gpurun_out
log=gpurun_out/sanitize_${tool}.log
timeout 1200 compute-sanitizer --tool "$tool" --print-limit 2000 --show-backtrace no \
    python -m pytest $tests -m gpu -q -x > "$log" 2>&1
# error sites by frequency, then pytest's and the sanitizer's summaries
grep -A1 "Uninitialized\|Invalid\|Host API\|hazard" "$log" | grep " at \|access by" |
    sed "s/+0x[0-9a-f]*//; s/at 0x[0-9a-f]*/at ADDR/" | sort | uniq -c | sort -rn | head -20
tail -4 "$log"
