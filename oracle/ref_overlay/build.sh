#!/usr/bin/env bash
# Overlay build of the UNMODIFIED reference host code (Scene, SceneGraph, SceneManager,
# ExampleScenes, TextureImporter, Resources, Core/*) from where it lies under the reference
# checkout, plus the product's host shim (path-tracing_b200/host).  Reference sources are reached
# through symlinks in oracle/_ref/overlay/ so that their quote-includes of "Application.h"
# resolve to our Vulkan-free stand-in.  Nothing is copied into the repository; all outputs go to
# oracle/_ref/ (git-ignored, travels to the GPU box).  Does NOT run the reference's CMake.
#
#   build.sh [REFERENCE_ROOT]      default /root/reference
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${1:-/root/reference}"
OUT="$REPO/oracle/_ref"
OV="$OUT/overlay"
OBJ="$OUT/obj"
PT="$REF/Path-Tracing"
HOST="$REPO/path-tracing_b200/host"
CXX="${CXX:-g++}"

if [ ! -d "$PT" ]; then
    echo "reference checkout not found at $REF: skipping overlay build" >&2
    exit 0
fi

mkdir -p "$OV/Core" "$OBJ"
link() { ln -sfn "$PT/$1" "$OV/$1"; }
for f in Scene.h Scene.cpp SceneGraph.h SceneGraph.cpp SceneManager.h SceneManager.cpp SceneImporter.h \
         ExampleScenes.h ExampleScenes.cpp TextureImporter.h TextureImporter.cpp Resources.h Resources.cpp \
         Core/Core.h Core/Core.cpp Core/Config.h Core/Config.cpp Core/Camera.h Core/Camera.cpp Core/Input.h \
         Core/Cache.h Core/Threads.h Shaders; do
    link "$f"
done
cp "$HERE/Application.h" "$OV/Application.h"

INC=(-I"$OV" -I"$REF/vendor/glm" -I"$REF/vendor/spdlog/include" -I"$REF/vendor/stb/stb" -I"$REF/vendor/gli"
     -I"$REF/vendor/glfw/include" -I"$REPO/include" -I"$HOST")
FLAGS=(-std=c++20 -O2 -fPIC -DCONFIG_PROFILE -DNDEBUG -w)

compile() { # src obj
    if [ ! -f "$2" ] || [ "$1" -nt "$2" ]; then
        echo "  CXX $(basename "$1")"
        "$CXX" "${FLAGS[@]}" "${INC[@]}" -c "$1" -o "$2"
    fi
}

# with the vendored assimp built (build_assimp.sh), the reference's own
# SceneImporter.cpp joins the overlay instead of the stub
ASSIMP_LIB="$OUT/assimp/libassimp.a"
EXTRA_LIBS=()
if [ -f "$ASSIMP_LIB" ]; then
    link SceneImporter.cpp
    INC+=(-I"$OUT/assimp/gen" -I"$REF/vendor/assimp/include")
    FLAGS+=(-DPT_OVERLAY_HAS_ASSIMP)
    EXTRA_LIBS=("$ASSIMP_LIB")
fi

REF_OBJS=()
if [ -f "$ASSIMP_LIB" ]; then
    compile "$OV/SceneImporter.cpp" "$OBJ/ref_SceneImporter.o"
    REF_OBJS+=("$OBJ/ref_SceneImporter.o")
fi
for f in Scene SceneGraph SceneManager ExampleScenes TextureImporter Resources Core/Core Core/Config Core/Camera; do
    o="$OBJ/ref_$(echo "$f" | tr / _).o"
    compile "$OV/$f.cpp" "$o"
    REF_OBJS+=("$o")
done
compile "$REF/vendor/stb/implementation.cpp" "$OBJ/stb_impl.o"
compile "$HERE/stubs.cpp" "$OBJ/stubs.o"
compile "$HOST/SceneFlatten.cpp" "$OBJ/SceneFlatten.o"
compile "$HOST/HeadlessRenderer.cpp" "$OBJ/HeadlessRenderer.o"
compile "$HERE/scene_dump.cpp" "$OBJ/scene_dump.o"
compile "$HERE/pt_headless.cpp" "$OBJ/pt_headless.o"

echo "  LD  scene_dump"
"$CXX" -o "$OUT/scene_dump" "$OBJ/scene_dump.o" "$OBJ/SceneFlatten.o" "$OBJ/stubs.o" "$OBJ/stb_impl.o" \
    "${REF_OBJS[@]}" "${EXTRA_LIBS[@]}" -pthread

CORE="$REPO/path-tracing_b200/csrc/libpt_core.so"
if [ -f "$CORE" ]; then
    echo "  LD  pt_headless"
    "$CXX" -o "$OUT/pt_headless" "$OBJ/pt_headless.o" "$OBJ/HeadlessRenderer.o" "$OBJ/SceneFlatten.o" \
        "$OBJ/stubs.o" "$OBJ/stb_impl.o" "${REF_OBJS[@]}" "${EXTRA_LIBS[@]}" "$CORE" -Wl,-rpath,'$ORIGIN/../../path-tracing_b200/csrc' -pthread
else
    echo "  (libpt_core.so not built yet: skipping pt_headless)"
fi
