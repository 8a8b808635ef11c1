#!/usr/bin/env bash
# Builds the assimp the reference vendors (vendor/assimp, an un-modified submodule) from where it
# lies, WITHOUT its CMake: glTF / glTF2 / OBJ importers only, no exporters, its own zlib.  The three
# headers CMake would generate (assimp/config.h, assimp/revision.h, zconf.h) are derived from their
# templates into oracle/_ref/assimp/gen/.  Output: oracle/_ref/assimp/libassimp.a (git-ignored).
# lets the overlay link the reference's own SceneImporter.cpp.
#
#   build_assimp.sh [REFERENCE_ROOT]      default /root/reference
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${1:-/root/reference}"
A="$REF/vendor/assimp"
OUT="$REPO/oracle/_ref/assimp"
GEN="$OUT/gen"
OBJ="$OUT/obj"
CXX="${CXX:-g++}"
CC="${CC:-gcc}"
JOBS="${JOBS:-$(nproc)}"

if [ ! -d "$A/code" ]; then
    echo "vendored assimp not found at $A: skipping" >&2
    exit 0
fi
if [ -f "$OUT/libassimp.a" ] && [ "$OUT/libassimp.a" -nt "$HERE/build_assimp.sh" ]; then
    echo "  (libassimp.a is up to date)"
    exit 0
fi
mkdir -p "$GEN/assimp" "$OBJ"

# generated headers
sed -e 's@^#cmakedefine \(.*\)$@/* #undef \1 */@' "$A/include/assimp/config.h.in" > "$GEN/assimp/config.h"
sed -e 's/@GIT_COMMIT_HASH@/0/g' -e 's/@GIT_BRANCH@/vendored/g' -e 's/@ASSIMP_VERSION_MAJOR@/6/g' \
    -e 's/@ASSIMP_VERSION_MINOR@/0/g' -e 's/@ASSIMP_VERSION_PATCH@/0/g' -e 's/@ASSIMP_PACKAGE_VERSION@/0/g' \
    -e 's/@[A-Za-z_]*@/0/g' "$A/include/assimp/revision.h.in" > "$GEN/assimp/revision.h"
sed -e 's@^#cmakedefine \(.*\)$@/* #undef \1 */@' "$A/contrib/zlib/zconf.h.cmakein" > "$GEN/zconf.h"

# every importer except glTF, glTF2 and OBJ is compiled out
DEFS=(-DASSIMP_BUILD_NO_EXPORT -DASSIMP_BUILD_NO_OWN_ZLIB_OFF -DRAPIDJSON_HAS_STDSTRING=1 -DRAPIDJSON_NOMEMBERITERATORCLASS
      -DASSIMP_BUILD_NO_M3D_EXPORTER -DOPENDDL_STATIC_LIBARY)
for imp in $(grep -o "ASSIMP_BUILD_NO_[A-Z0-9_]*_IMPORTER" "$A/code/Common/ImporterRegistry.cpp" | sort -u); do
    case "$imp" in
    ASSIMP_BUILD_NO_GLTF_IMPORTER | ASSIMP_BUILD_NO_GLTF1_IMPORTER | ASSIMP_BUILD_NO_GLTF2_IMPORTER | ASSIMP_BUILD_NO_OBJ_IMPORTER) ;;
    *) DEFS+=("-D$imp") ;;
    esac
done
INC=(-I"$GEN" -I"$A/include" -I"$A" -I"$A/code" -I"$A/contrib" -I"$A/contrib/zlib" -I"$A/contrib/unzip" -I"$A/contrib/rapidjson/include"
     -I"$A/contrib/utf8cpp/source" -I"$A/contrib/pugixml/src" -I"$A/contrib/openddlparser/include" -I"$A/contrib/stb")

SRCS=()
for d in Common CApi PostProcessing Material Geometry AssetLib/glTF AssetLib/glTF2 AssetLib/glTFCommon AssetLib/Obj; do
    for f in "$A/code/$d"/*.cpp; do
        [ -f "$f" ] && SRCS+=("$f")
    done
done
CSRCS=("$A"/contrib/zlib/*.c "$A/contrib/unzip/unzip.c" "$A/contrib/unzip/ioapi.c")

export A OBJ CC CXX
# arrays do not survive export: pass them through the environment as strings
export DEFS_STR="${DEFS[*]}" INC_STR="${INC[*]}"
printf '%s\n' "${SRCS[@]}" "${CSRCS[@]}" | xargs -P "$JOBS" -I{} bash -c '
    src="{}"; o="$OBJ/$(echo "${src#$A/}" | tr / _).o"
    if [ ! -f "$o" ] || [ "$src" -nt "$o" ]; then
        case "$src" in
        *.c) $CC -O1 -fPIC -w $INC_STR -c "$src" -o "$o" ;;
        *) $CXX -std=c++17 -O1 -fPIC -w $DEFS_STR $INC_STR -c "$src" -o "$o" ;;
        esac || { echo "FAILED: $src" >&2; exit 255; }
    fi'
rm -f "$OUT/libassimp.a"
ar rcs "$OUT/libassimp.a" "$OBJ"/*.o
echo "  AR  libassimp.a ($(ls "$OBJ" | wc -l) objects)"
