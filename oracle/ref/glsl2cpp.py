#!/usr/bin/env python3
"""ORACLE -- TEST INFRASTRUCTURE ONLY.

glsl2cpp.py: mechanical rewrite of a reference shader (read from /root/reference, never copied into this repository)
into text g++ accepts on top of oracle/ref/glsl_shim.h.  The output goes to oracle/_ref/gen/ (git-ignored).  The rules
are purely syntactic -- no arithmetic is added, removed or reordered:

  1. `#include "x"` is inlined (what the host's ReadWithPreprocessor does, src/Base/src/Utils.cpp); `#version` /
     `#extension` lines are dropped; include guards and `#if` permutations are left to the C preprocessor.
  2. every real literal gets an `f` suffix (GLSL literals are fp32; C++'s would be double).
  3. `layout(...)` qualifiers and the `uniform` / `shared` / `readonly` / `writeonly` / `restrict` / `coherent` keywords are
     dropped: uniform-block and storage-block members, samplers, images and shared arrays become namespace-scope variables the driver fills;
     `layout(local_size...) in;` lines disappear (the driver passes the local size to ref::dispatch).
  4. parameter qualifiers: `out T x` / `inout T x` -> `T& x`, `in T x` -> `T x`; file-scope `in` / `out` interface
     variables of a fragment shader become thread_local variables (one fragment per thread).
  5. GLSL array constructors `T[](...)` / `T[n](...)` -> `{...}`; `discard` -> `REF_DISCARD` (a macro: `return`, or the helper-invocation flag of a
     program that takes derivatives).
  6. `imageSize(x)` / `textureSize(x, l)` keep their names (the shim overloads them on the image / sampler type).
  8. GLSL-style array types `T[N] name` (function return types, locals) become std::array<T, N>.
  7. file-scope scalar / vector variables with an initialiser become macros (GLSL initialises them per invocation,
     after the uniforms are bound; a C++ global would be initialised once, at load time).

usage: glsl2cpp.py <shader path> <output path> [search dir ...]
"""
import os
import re
import sys


def inline_includes(path, search, seen_depth=0):
    if seen_depth > 16:
        raise RuntimeError("include depth")
    out = []
    here = os.path.dirname(path)
    with open(path, encoding="latin-1") as f:
        for line in f:
            m = re.match(r'\s*#include\s+"([^"]+)"', line)
            if m:
                name = m.group(1)
                for d in [here] + search:
                    cand = os.path.normpath(os.path.join(d, name))
                    if os.path.exists(cand):
                        out.append(f"// ---- #include \"{name}\"\n")
                        out.extend(inline_includes(cand, search, seen_depth + 1))
                        out.append(f"\n// ---- end of \"{name}\"\n")
                        break
                else:
                    raise FileNotFoundError(f"{name} included from {path}")
            elif re.match(r"\s*#(version|extension)\b", line):
                continue
            else:
                out.append(line)
    if out and not out[-1].endswith("\n"):
        out[-1] += "\n"
    return out


REAL = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def rewrite(text):
    # strip comments first (non-ASCII comments in the reference would otherwise confuse nothing, but keep output small)
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    # 2. real literals
    text = REAL.sub(lambda m: m.group(1) + "f", text)
    # 3. layout(...) in;  -> gone
    text = re.sub(r"layout\s*\([^)]*\)\s*in\s*;", "", text)
    # uniform blocks: layout(std140, binding = N) uniform Name { members } [instance];  -> members at namespace scope
    def block(m):
        body, inst = m.group(2), m.group(3)
        if inst:
            return f"struct {m.group(1)}_t {{{body}}} {inst};"
        return body
    text = re.sub(r"layout\s*\([^)]*\)\s*(?:uniform|buffer)\s+(\w+)\s*\{(.*?)\}\s*(\w*)\s*;", block, text, flags=re.S)
    text = re.sub(r"layout\s*\([^)]*\)\s*", "", text)
    text = re.sub(r"\b(uniform|shared|readonly|writeonly|restrict|coherent|highp|mediump|lowp|flat)\s+", "", text)
    # 4. file-scope interface variables of fragment shaders
    text = re.sub(r"^(\s*)(in|out)\s+(\w+\s+\w+\s*;)", r"\1thread_local \3", text, flags=re.M)
    # parameter qualifiers
    text = re.sub(r"\b(?:out|inout)\s+(\w+)\s+(\w+)", r"\1& \2", text)
    text = re.sub(r"([(,]\s*(?:const\s+)?)in\s+(\w+\s+\w+)", r"\1\2", text)
    # 5. array constructors and discard
    text = re.sub(r"=\s*\w+\s*\[\s*\w*\s*\]\s*\(", "= REF_ARRAY_BEGIN(", text)
    text = convert_array_ctors(text)
    text = re.sub(r"\bdiscard\s*;", "REF_DISCARD;", text)
    # 8. array types written GLSL-style: `T[N] name` (return types, locals) -> std::array<T, N> name
    text = re.sub(r"\b([A-Za-z_]\w*)\s*\[\s*(\d+)\s*\]\s+([A-Za-z_]\w*)", r"std::array<\1, \2> \3", text)
    # 7. file-scope scalars / vectors with initialisers (`const vec3 kCloudAABBMin = vec3(..., uBottomAltitude);`): GLSL
    #    evaluates them per invocation, after the uniforms are set -- C++ would at load time.  They become macros.
    text = globals_to_macros(text)
    return text


GLOBAL_INIT = re.compile(r"(?:const\s+)?\b(float|int|uint|bool|vec[234]|ivec[234]|uvec[234])\s+(\w+)\s*=\s*([^;{}]+);")


def globals_to_macros(text):
    out, depth, i = [], 0, 0
    while i < len(text):
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
        if depth == 0 and (i == 0 or text[i - 1] in "\n;}"):
            m = GLOBAL_INIT.match(text, i + (1 if c in " \t" else 0)) if c in " \tcfiubv" else None
            if m and text[i:m.start()].strip() == "":
                rhs = " ".join(m.group(3).split())
                out.append(f"\n#define {m.group(2)} ({m.group(1)}({rhs}))\n")
                i = m.end()
                continue
        out.append(c)
        i += 1
    return "".join(out)


def convert_array_ctors(text):
    """`= REF_ARRAY_BEGIN( a, b, c );` -> `= { a, b, c };` (balanced parentheses)."""
    key = "REF_ARRAY_BEGIN("
    while True:
        i = text.find(key)
        if i < 0:
            return text
        j = i + len(key)
        depth = 1
        while depth:
            c = text[j]
            depth += c == "("
            depth -= c == ")"
            j += 1
        text = text[:i] + "{" + text[i + len(key):j - 1] + "}" + text[j:]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    search = sys.argv[3:]
    text = rewrite("".join(inline_includes(src, search)))
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        f.write(f"// generated by oracle/ref/glsl2cpp.py from {src} -- do not commit\n")
        f.write(text)


if __name__ == "__main__":
    main()
