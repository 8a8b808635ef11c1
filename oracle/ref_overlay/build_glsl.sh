#!/usr/bin/env bash
# "The reference compiled here" for the shader half of the hot path: turns the reference's GLSL
# ray-tracing stages (raygen.rgen, closestHit.rchit, anyhit.rahit, occlusionAnyhit.rahit, miss.rmiss,
# occlusion.rmiss and everything they #include) into C++ by the mechanical transform glsl2cpp.py —
# reading the sources where they lie under the reference checkout — and compiles the result against
# the reference's own vendored glm into oracle/_ref/libglsl_ref.so.  The generated C++ goes to
# oracle/_ref/glsl/ (git-ignored): no reference source enters the repository.
#
#   build_glsl.sh [REFERENCE_ROOT]      default /root/reference
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${1:-/root/reference}"
OUT="$REPO/oracle/_ref"
CXX="${CXX:-g++}"

if [ ! -d "$REF/Path-Tracing/Shaders" ]; then
    echo "reference checkout not found at $REF: skipping GLSL build" >&2
    exit 0
fi
mkdir -p "$OUT/glsl"
GEN="$OUT/glsl/glsl_stages.cpp"
LIB="$OUT/libglsl_ref.so"
CGEN="$OUT/glsl/glsl_compute.cpp"
CLIB="$OUT/libglsl_comp_ref.so"
newest=$(ls -t "$HERE"/glsl2cpp.py "$HERE"/glsl_rt*.h "$REPO/include/pt_core.h" "$REF"/Path-Tracing/Shaders/*.* | head -1)
if [ -f "$LIB" ] && [ "$LIB" -nt "$newest" ] && [ -f "$CLIB" ] && [ "$CLIB" -nt "$newest" ]; then
    exit 0
fi
python3 "$HERE/glsl2cpp.py" "$REF" "$GEN"
echo "  CXX glsl_stages.cpp -> libglsl_ref.so"
# -ffp-contract=off: GLSL fuses only where the source says fma(); -O2 without fast-math keeps IEEE semantics
"$CXX" -std=c++20 -O2 -fPIC -shared -ffp-contract=off -fno-fast-math -fvisibility=hidden -w \
    -I"$HERE" -I"$REF/vendor/glm" -o "$LIB" "$GEN" -pthread
# the compute stages (post-process chain, skinning): same transform, their own runtime headers
python3 "$HERE/glsl2cpp.py" --compute "$REF" "$CGEN"
echo "  CXX glsl_compute.cpp -> libglsl_comp_ref.so"
"$CXX" -std=c++20 -O2 -fPIC -shared -ffp-contract=off -fno-fast-math -fvisibility=hidden -w \
    -I"$HERE" -I"$REF/vendor/glm" -o "$CLIB" "$CGEN" -pthread
