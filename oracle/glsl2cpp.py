#!/usr/bin/env python3
"""Build-time translator (test infrastructure): turns a reference GLSL shader, read where it lies under /root/reference, into a
C++ translation unit that oracle/ref_shader_driver.cpp compiles against the reference's vendored glm (oracle/glsl_compat.h supplies
samplers, images and the int/float mixed operators GLSL has and glm lacks).  Nothing is copied into the repo: the output goes to
oracle/_ref/, which is git-ignored.  Only syntax is rewritten — declarations the C++ compiler cannot read; every expression of the
shader is compiled as written.

  usage: glsl2cpp.py <shader.glsl> <namespace> <out.inc> [extra,pinned,functions]
"""
import re
import sys


def match_paren(s, i):
    """index of the parenthesis closing the one at s[i]"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses")


USED_SWIZZLES = set()
SWZ_OPS = {"=": "'='", "+=": "'+'", "-=": "'-'", "*=": "'*'", "/=": "'/'"}


def rewrite_swizzles(s):
    pat = re.compile(r"\.([xyzw]{2,4}|[rgba]{2,4})\b")
    pos = 0
    while True:
        m = pat.search(s, pos)
        if not m:
            return s
        # walk left over the postfix expression the swizzle applies to: identifiers, member chains, calls and subscripts
        i = m.start()
        while i > 0:
            c = s[i - 1]
            if c in ")]":
                close, opener = c, "(" if c == ")" else "["
                depth, j = 0, i - 1
                while j >= 0:
                    if s[j] == close:
                        depth += 1
                    elif s[j] == opener:
                        depth -= 1
                        if depth == 0:
                            break
                    j -= 1
                i = j
            elif c.isalnum() or c == "_":
                while i > 0 and (s[i - 1].isalnum() or s[i - 1] == "_"):
                    i -= 1
                if i > 0 and s[i - 1] == ".":
                    i -= 1
                    continue
                break
            else:
                break
        expr = s[i:m.start()]
        USED_SWIZZLES.add(m.group(1))
        rep = f"swz_{m.group(1)}({expr})"
        s = s[:i] + rep + s[m.end():]
        pos = i + len(rep)


def translate(src, ns, extra_pinned=()):
    out = []
    lines = src.replace("\r\n", "\n").split("\n")
    k = 0
    while k < len(lines):
        ln = lines[k]
        st = ln.strip()
        if st.startswith("#version") or st.startswith("#extension"):
            k += 1
            continue
        # layout(local_size_x = ...) in;   (compute work-group size: the driver loops over invocations itself)
        if re.match(r"layout\s*\(\s*local_size", st):
            k += 1
            continue
        # layout(std430, binding = N) buffer Name { members };  ->  the members become globals of the shader's namespace
        m = re.match(r"layout\s*\(\s*std430.*\)\s*buffer\s+(\w+)", st)
        if m:
            # the members of one block are contiguous in the buffer object: emit one array and a pointer per member, so that a
            # shader indexing past a member's end reads what the GL buffer would hold there (the driver fills the tail padding)
            block = m.group(1)
            k += 1
            while "{" not in lines[k - 1] and "{" not in lines[k]:
                k += 1
            if "{" in lines[k]:
                k += 1
            members = []
            while not lines[k].strip().startswith("}"):
                mm = re.match(r"\s*(\w+)\s+(\w+)\s*\[([^\]]*)\]\s*;", lines[k])
                if mm:
                    members.append((mm.group(1), mm.group(2), mm.group(3).strip() or "1"))
                else:
                    mm = re.match(r"\s*(\w+)\s+(\w+)\s*;", lines[k])   # scalar member (uint SkyLevelAggregate;)
                    if mm:
                        members.append((mm.group(1), mm.group(2), None))
                k += 1
            k += 1
            types = {t for t, _, _ in members}
            if any(n is None for _, _, n in members):
                for t, name, n in members:
                    out.append(f"static {t} {name};" if n is None else f"static {t} {name}[({n}) + 64];")
            elif len(types) == 1:
                t = members[0][0]
                total = " + ".join(f"({n})" for _, _, n in members)
                out.append(f"static {t} {block}_storage[{total} + 4096];")
                off = "0"
                for _, name, n in members:
                    out.append(f"static {t}* const {name} = {block}_storage + ({off});")
                    off = f"{off} + ({n})"
            else:
                for t, name, n in members:
                    out.append(f"static {t} {name}[({n}) + 64];")
            continue
        # layout(r8, binding = 0) uniform image3D name;   /   layout (location = N) out float name;
        ln = re.sub(r"^\s*layout\s*\([^)]*\)\s*uniform\s+", "static ", ln)
        ln = re.sub(r"^\s*layout\s*\([^)]*\)\s*out\s+", "static ", ln)
        ln = re.sub(r"^\s*uniform\s+", "static ", ln)
        ln = re.sub(r"^\s*in\s+(vec\d|float|int)\s+", r"static \1 ", ln)
        out.append(ln)
        k += 1
    s = "\n".join(out)
    s = re.sub(r"(\w+)\s*\[\s*\]\s*;", r"\1[1];", s)
    # parameter qualifiers
    s = re.sub(r"\bconst\s+in\s+", "const ", s)
    s = re.sub(r"([(,]\s*)(?:inout|out)\s+(\w+)\s+(\w+)", r"\1\2& \3", s)
    s = re.sub(r"([(,]\s*)in\s+(\w+)\s+(\w+)", r"\1\2 \3", s)
    # arrays as values:  float[6] f(...)  /  float[6] x = ...;  /  float x[6] = f(...);   ->  farr<6>
    s = re.sub(r"\bfloat\s*\[\s*(\d+)\s*\]\s+(\w+)", r"farr<\1> \2", s)
    s = re.sub(r"\bfloat\s+(\w+)\s*\[\s*(\d+)\s*\]\s*=\s*(\w+)\s*\(", r"farr<\2> \1 = \3(", s)
    s = re.sub(r"(?m)^(\s*)float\s+(\w+)\s*\[\s*(\d+)\s*\]\s*;", r"\1farr<\3> \2;", s)
    # array constructors:  const vec3 N[6] = vec3[]( a, b, c );  ->  const vec3 N[6] = { a, b, c };
    while True:
        m = re.search(r"\b(i?vec[234]|float|int|farr<\d+>)\s*(?:\[\s*\d*\s*\])?\s*\(", s) if False else re.search(r"\b(?:(?:i?vec[234]|float|int)\s*\[\s*\d*\s*\]|farr<\d+>)\s*\(", s)
        if not m:
            break
        close = match_paren(s, m.end() - 1)
        s = s[:m.start()] + "{" + s[m.end():close] + "}" + s[close + 1:]
    # floating literals: GLSL's unsuffixed literals are fp32 (C++'s would be double and change every expression they touch)
    s = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])", r"\1f", s)
    # swizzle statements used as l-values:  X.xy = e;  X.xyz /= e;
    s = re.sub(r"(?m)^(\s*)(\w+)\.(xy|xyz|rgb)\s*(=|\+=|-=|\*=|/=)\s*([^;=][^;]*);", lambda m: f"{m.group(1)}swz_assign_{m.group(3)}({m.group(2)}, {SWZ_OPS[m.group(4)]}, {m.group(5)});", s)
    # r-value swizzles glm does not provide as members:  <postfix-expression>.xyz  ->  swz_xyz(<postfix-expression>)
    s = rewrite_swizzles(s)
    # transcendental functions: pinned definitions (glsl_compat.h)
    s = re.sub(r"\b(sin|cos|tan|pow|mix|log2|acos)\s*\(", r"pinned_\1(", s)
    for fn in extra_pinned:   # per-shader additions (the denoising passes: exp), so that shaders translated earlier keep their text
        s = re.sub(r"\b(%s)\s*\(" % re.escape(fn), r"pinned_\1(", s)
    # GLSL evaluates function arguments left to right (spec 6.1.1); C++ leaves the order of constructor arguments open (g++: right to
    # left).  Constructor calls whose arguments advance the hash RNG are brace-initialised, which C++ orders left to right.
    s = re.sub(r"\bvec2\s*\(\s*(Hash1?\(\))\s*,\s*(Hash1?\(\))\s*\)", r"vec2{\1, \2}", s)
    # entry point
    s = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", s)
    # GLSL re-initialises global variables for every shader invocation: collect the initialised, mutable globals (brace depth 0)
    # and emit shader_reset_globals(), which the driver calls before each shader_main()
    resets, depth = [], 0
    for ln in s.split("\n"):
        if depth == 0:
            m = re.match(r"^(?:int|uint|float|bool|i?vec[234]|mat[234])\s+(\w+)\s*=\s*(.+);\s*(?://.*)?$", ln.strip())
            if m:
                resets.append(f"    {m.group(1)} = {m.group(2)};")
        depth += ln.count("{") - ln.count("}")
    s += "\ninline void shader_reset_globals() {\n" + "\n".join(resets) + "\n}\n"
    helpers = []
    for name in sorted(USED_SWIZZLES):
        comps = ", ".join("v." + "xyzw"["xyzw".index(c) if c in "xyzw" else "rgba".index(c)] for c in name)
        helpers.append(f"template <class V> inline vec{len(name)} swz_{name}(const V& v) {{ return vec{len(name)}({comps}); }}")
    USED_SWIZZLES.clear()
    return f"namespace {ns} {{\n" + "\n".join(helpers) + f"\n{s}\n}}  // namespace {ns}\n"


if __name__ == "__main__":
    path, ns, dst = sys.argv[1:4]
    extra = sys.argv[4].split(",") if len(sys.argv) > 4 else ()
    with open(path, encoding="utf-8", errors="replace") as f:
        text = translate(f.read(), ns, extra)
    with open(dst, "w") as f:
        f.write(f"// generated at build time from {path} by oracle/glsl2cpp.py — do not commit\n" + text)
