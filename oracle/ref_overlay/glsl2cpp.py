#!/usr/bin/env python3
"""glsl2cpp.py — mechanical GLSL -> C++ transform of the reference's ray-tracing shaders.

TEST INFRASTRUCTURE (oracle/_ref: "the reference compiled here").  Reads the shader sources where
they lie under <reference>/Path-Tracing/Shaders and writes ONE generated C++ translation unit (into
oracle/_ref/, which is git-ignored: no reference source enters the repository).  The generated file
is compiled against the glm the reference vendors, so that every arithmetic expression of the GLSL is
evaluated by a C++ compiler exactly as written.

The transform is purely textual; nothing is re-derived or re-ordered:

  1. `#include "x"` is expanded in place, once per stage (the reference's shaderc includer does the
     same: Renderer/ShaderLibrary.cpp:416 "#pragma once"); the three `.incl` type headers are emitted
     once, in front of all stages, through their GLSL (`GL_core_profile`) branch.
  2. `#version`, `#extension` lines and every `layout(...) ... ;` declaration (descriptor bindings,
     payload / hit-attribute / shader-record / specialisation-constant declarations) are dropped; the
     runtime header glsl_rt.h declares the same names (payload, textures, transforms, geometries, the
     material buffers, the light block, sbt, attribs, s_HitFlags, s_MissFlags, gl_* built-ins).
  3. buffer references: `layout(buffer_reference ...) buffer Name { Type[] v; }` -> `struct Name { Type *v; }`.
  4. parameter qualifiers: `out T x` / `inout T x` -> `T &x`; `in T x` -> `T x`.
  5. unsuffixed floating-point literals get an `f` suffix (a GLSL `1.0` is a 32-bit float, a C++ `1.0`
     is a double).
  6. swizzles `.xyz` -> `.xyz()` (glm's function swizzles, GLM_FORCE_SWIZZLE on gcc); single
     components are plain members in both languages.
  7. `void main()` -> `void main_()`, `ignoreIntersectionEXT;` -> `{ glsl_ignore_intersection(); return; }`.
  8. a vector constructor whose arguments call rand() more than once is written with braces,
     `vec2(rand(s), rand(s))` -> `vec2{rand(s), rand(s)}`: GLSL evaluates arguments left to right
     (GLSL 4.60 §6.1.1), C++ only does so inside a braced initialiser list (g++ goes right to left).
     The debug stages' `payload` (a DebugPayload) is mapped to its own object by a #define around those stages.
  9. functions listed in DROP_FUNCTIONS (GLSL array constructors; skinning helpers that are not on the
     ray-tracing path) are removed.

  10. (compute stages only) a GLSL array constructor assigned to an array member,
     `v.BoneIndices = uint[4](a, b, c, d);`, is written element by element:
     `{ v.BoneIndices[0] = (a); ... v.BoneIndices[3] = (d); }` (C++ has no array assignment).

With --compute the COMPUTE stages (postprocess.comp, bloomDownsample.comp, bloomUpsample.comp, composition.comp,
toneMapping.comp, skinning.comp) are generated instead, each in its own namespace behind the per-stage state macro of
glsl_rt_compute.h (rows f2 / f3 of SURVEY.md 8).

Usage: glsl2cpp.py [--compute] <reference root> <output .cpp>
"""
import os
import re
import sys

STAGES = [
    # namespace, file (relative to Path-Tracing/Shaders)
    ("rgen", "raygen.rgen"),
    ("rchit", "closestHit.rchit"),
    ("rahit", "anyhit.rahit"),
    ("occ_rahit", "occlusionAnyhit.rahit"),
    ("rmiss", "miss.rmiss"),
    ("occ_rmiss", "occlusion.rmiss"),
    # the debug pipeline (Renderer.cpp:579-589): its own raygen / miss / hit group; occlusion rays use the two above
    ("dbg_rgen", "Debug/debugRaygen.rgen"),
    ("dbg_rchit", "Debug/debugClosestHit.rchit"),
    ("dbg_rahit", "Debug/debugAnyhit.rahit"),
    ("dbg_rmiss", "Debug/debugMiss.rmiss"),
]
COMPUTE_STAGES = [
    ("comp_post", "postprocess.comp"),
    ("comp_down", "bloomDownsample.comp"),
    ("comp_up", "bloomUpsample.comp"),
    ("comp_compose", "composition.comp"),
    ("comp_tone", "toneMapping.comp"),
    ("comp_skin", "skinning.comp"),
]
TYPE_HEADERS = ["ShaderTypes.incl", "ShaderRendererTypes.incl", "Debug/DebugShaderTypes.incl"]
DROP_FUNCTIONS = ["getAnimatedVertex", "writeVertex"]


def strip_preprocessor_branches(text: str) -> str:
    """Evaluates #ifndef GL_core_profile / #ifdef / #else / #endif with GL_core_profile DEFINED (the
    GLSL branch of the dual headers) and drops #pragma once / #define BUFFER_POINTER (rule 3)."""
    out = []
    stack = []  # True = emitting
    for line in text.splitlines():
        s = line.strip()
        if s.startswith("#ifndef GL_core_profile"):
            stack.append(False)
            continue
        if s.startswith("#ifdef GL_core_profile"):
            stack.append(True)
            continue
        if s.startswith("#else") and stack:
            stack[-1] = not stack[-1]
            continue
        if s.startswith("#endif") and stack:
            stack.pop()
            continue
        if all(stack):
            out.append(line)
    return "\n".join(out) + "\n"


def common_rules(text: str) -> str:
    # rule 3: the buffer-reference macro of ShaderRendererTypes.incl
    text = re.sub(
        r"#define\s+BUFFER_POINTER\(Name,\s*Type\)\s+layout\(buffer_reference[^)]*\)\s*buffer\s+Name\s*\{\s*Type\[\]\s*v;\s*\}",
        "#define BUFFER_POINTER(Name, Type) struct Name { Type *v; }",
        text,
    )
    # rule 2
    text = re.sub(r"^\s*#(version|extension)[^\n]*\n", "", text, flags=re.M)
    text = re.sub(r"layout\s*\([^)]*\)[^;{]*(\{[^}]*\}\s*)?[^;{]*;", "", text, flags=re.S)
    text = re.sub(r"^\s*hitAttributeEXT[^;]*;", "", text, flags=re.M)
    # rule 4
    text = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1 &\2", text)
    text = re.sub(r"([(,]\s*)in\s+(\w+\s+\w+)", r"\1\2", text)
    # rule 5 (not inside identifiers, not already suffixed, not hex)
    text = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])", r"\1f", text)
    # rule 6
    text = re.sub(r"\.([xyzw]{2,4}|[rgba]{2,4}|[stpq]{2,4})\b(?!\s*\()", r".\1()", text)
    # rule 7
    text = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void main_()", text)
    text = re.sub(r"\bignoreIntersectionEXT\s*;", "{ glsl_ignore_intersection(); return; }", text)
    return order_side_effects(text)


def order_side_effects(text: str) -> str:
    """rule 8"""
    out = []
    i = 0
    pat = re.compile(r"\bvec[234]\(")
    while True:
        m = pat.search(text, i)
        if not m:
            out.append(text[i:])
            break
        j = m.end()
        depth = 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        inner = text[m.end() : j - 1]
        if inner.count("rand(") >= 2:
            out.append(text[i : m.end() - 1] + "{" + inner + "}")
            i = j
        else:
            out.append(text[i : m.end()])
            i = m.end()
    return "".join(out)


def drop_function(text: str, name: str) -> str:
    """Removes `<type> name(...) { ... }` (brace matched); rule 9."""
    m = re.search(r"^[\w\[\]]+\s+" + re.escape(name) + r"\s*\(", text, flags=re.M)
    if not m:
        return text
    i = text.index("{", m.end())
    depth = 0
    j = i
    while True:
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return text[: m.start()] + f"/* {name}: dropped (rule 9) */" + text[j + 1 :]


def array_constructors(text: str) -> str:
    """rule 10"""

    def repl(m):
        target, count, args = m.group(1), int(m.group(3)), m.group(4)
        parts, depth, cur = [], 0, ""
        for ch in args:
            if ch == "," and depth == 0:
                parts.append(cur.strip())
                cur = ""
                continue
            depth += {"(": 1, ")": -1}.get(ch, 0)
            cur += ch
        parts.append(cur.strip())
        assert len(parts) == count, (target, parts)
        return "{ " + " ".join(f"{target}[{i}] = ({a});" for i, a in enumerate(parts)) + " }"

    return re.sub(r"([\w.]+)\s*=\s*(uint|int|float)\[(\d+)\]\((.*)\);", repl, text)


def expand(path: str, shader_dir: str, seen: set, skip: set) -> str:
    rel = os.path.relpath(path, shader_dir)
    if rel in seen or rel in skip:
        return f"/* #include \"{rel}\": already included */\n"
    seen.add(rel)
    with open(path, "r", encoding="utf-8-sig") as f:
        src = f.read()
    out = [f"/* ---- begin {rel} ---- */\n#line 1 \"{path}\"\n"]
    for n, line in enumerate(src.splitlines(), 1):
        m = re.match(r'\s*#include\s+"([^"]+)"', line)
        if m:
            inc = os.path.normpath(os.path.join(os.path.dirname(path), m.group(1)))
            if not os.path.exists(inc):
                inc = os.path.normpath(os.path.join(shader_dir, m.group(1)))
            out.append(expand(inc, shader_dir, seen, skip))
            out.append(f"#line {n + 1} \"{path}\"\n")
        else:
            out.append(line + "\n")
    out.append(f"/* ---- end {rel} ---- */\n")
    return "".join(out)


def main():
    args = sys.argv[1:]
    compute = "--compute" in args
    if compute:
        args.remove("--compute")
    ref, out_path = args[0], args[1]
    shader_dir = os.path.join(ref, "Path-Tracing", "Shaders")
    parts = [
        "/* GENERATED by oracle/ref_overlay/glsl2cpp.py from the reference's shader sources — do not commit. */\n",
        '#include "glsl_rt.h"\n',
        "namespace glslref\n{\n",
    ]
    seen = set()
    types = "".join(expand(os.path.join(shader_dir, h), shader_dir, seen, set()) for h in TYPE_HEADERS)
    parts.append(common_rules(strip_preprocessor_branches(types)))
    state = "glsl_rt_compute.h" if compute else "glsl_rt_state.h"
    parts.append(f'\n}} // namespace glslref\n#include "{state}"\nnamespace glslref\n{{\n')
    skip = set(TYPE_HEADERS)
    for ns, fname in COMPUTE_STAGES if compute else STAGES:
        body = expand(os.path.join(shader_dir, fname), shader_dir, set(), skip)
        body = common_rules(strip_preprocessor_branches(body))
        if compute:
            body = array_constructors(body)
            parts.append(f"\nnamespace {ns}\n{{\nGLSL_COMPUTE_STATE_{ns}\n{body}\n}} // namespace {ns}\n")
            continue
        for fn in DROP_FUNCTIONS:
            body = drop_function(body, fn)
        if ns.startswith("dbg_"):
            # the debug stages declare `payload` as a DebugPayload (Debug/debugRaygen.rgen:18): same name, other object
            body = "#define payload dbg_payload\n" + body + "\n#undef payload\n"
        parts.append(f"\nnamespace {ns}\n{{\n{body}\n}} // namespace {ns}\n")
    api = "glsl_rt_compute_api.h" if compute else "glsl_rt_api.h"
    parts.append(f'\n}} // namespace glslref\n#include "{api}"\n')
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        f.write("".join(parts))


if __name__ == "__main__":
    main()
