"""Mechanical models of the C ABI (include/ibu_b200.h) and of its Rust binding (rust/src/gpu/ffi.rs).

No Rust toolchain exists in the build image, so the binding cannot be compiled here; instead both
files are parsed into the same shape — constants, structs (field names, types, order), opaque
handles, the callback type and every function signature — and tests/test_rust_ffi.py fails on any
difference.  rust/gen_ffi.py renders ffi.rs from the header model (run it after changing the header).
"""
from __future__ import annotations

import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ibu_b200.h")
FFI_RS = os.path.join(ROOT, "rust", "src", "gpu", "ffi.rs")

BASE = {"void": "c_void", "char": "c_char", "int": "c_int", "unsigned": "c_uint", "unsigned int": "c_uint",
        "size_t": "usize", "uint8_t": "u8", "uint32_t": "u32", "int32_t": "i32", "uint64_t": "u64", "double": "f64"}


def c_type_to_rust(ctype: str) -> str:
    """`const ibu_record_t *const *` -> `*const *const ibu_record_t`."""
    t = ctype.strip()
    levels = []  # own constness of each pointer level, OUTERMOST first (peeled from the right)
    while True:
        m = re.match(r"^(.*)\*\s*(const)?\s*$", t)
        if not m:
            break
        levels.append(bool(m.group(2)))
        t = m.group(1).strip()
    const_base = bool(re.search(r"\bconst\b", t))
    base = re.sub(r"\bconst\b", "", t).strip()
    base = re.sub(r"^struct\s+", "", base)
    rust = BASE.get(base, base)
    if not levels:
        return rust
    # the innermost pointer's pointee constness is the base's; each outer level's pointee constness
    # is the `const` written after the inner `*`
    inner_first = levels[::-1]
    pointee_const = [const_base] + inner_first[:-1]
    for pc in pointee_const:
        rust = ("*const " if pc else "*mut ") + rust
    return rust


def _strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def _num(text: str) -> int:
    expr = re.sub(r"(?<=[0-9a-fA-F])(ull|ULL|ul|UL|u|U|l|L)\b", "", text.strip())
    return int(eval(expr, {"__builtins__": {}}, {}))  # literals and * only


def _split_args(args: str):
    out, depth, cur = [], 0, ""
    for ch in args:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [a.strip() for a in out]


def _decl(arg: str):
    """`const ibu_record_t *d_records` -> (rust type, name); arrays `uint8_t reserved[8]` -> [u8; 8]."""
    arg = arg.strip()
    m = re.match(r"^(.*?)(\w+)\s*\[(\w+)\]$", arg)
    if m:
        return f"[{c_type_to_rust(m.group(1))}; {int(m.group(3))}]", m.group(2)
    m = re.match(r"^(.*?)(\w+)$", arg, flags=re.S)
    return c_type_to_rust(m.group(1)), m.group(2)


def parse_header(path: str = HEADER) -> dict:
    raw = open(path).read()
    src = _strip_comments(raw)
    model = {"consts": {}, "structs": {}, "opaque": [], "callbacks": {}, "fns": {}}
    for m in re.finditer(r"^[ \t]*#define[ \t]+(\w+)[ \t]+(.+?)[ \t]*$", src, flags=re.M):
        name, val = m.group(1), m.group(2)
        if name == "IBU_B200_H":
            continue
        try:
            model["consts"][name] = _num(val)
        except Exception:
            pass  # function-like or non-numeric macros are not part of the binding
    for m in re.finditer(r"typedef\s+enum\s+\w*\s*\{(.*?)\}\s*\w+\s*;|(?<!typedef )enum\s*\{(.*?)\}\s*;", src, flags=re.S):
        body = m.group(1) or m.group(2)
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = item.split("=")
                nxt = _num(v)
                model["consts"][k.strip()] = nxt
            else:
                model["consts"][item] = nxt
            nxt += 1
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            if decl.strip():
                t, n = _decl(decl)
                fields.append((n, t))
        model["structs"][m.group(3)] = fields
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s+(\w+)\s*;", src):
        model["opaque"].append(m.group(2))
    for m in re.finditer(r"typedef\s+(\w[\w\s\*]*?)\(\s*\*\s*(\w+)\s*\)\s*\((.*?)\)\s*;", src, flags=re.S):
        model["callbacks"][m.group(2)] = (c_type_to_rust(m.group(1)), [_decl(a) for a in _split_args(m.group(3))])
    body = re.sub(r"typedef\s+(struct|enum)\s+\w*\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)
    body = re.sub(r"(?<!typedef )enum\s*\{.*?\}\s*;", " ", body, flags=re.S)
    body = re.sub(r"typedef[^;]*;", " ", body)
    body = re.sub(r"^[ \t]*#.*$", " ", body, flags=re.M)
    body = body.replace('extern "C" {', " ")
    for m in re.finditer(r"([\w\s\*]+?)\b(ibu_\w+)\s*\(([^;{}]*?)\)\s*;", body, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = [] if args in ("void", "") else [_decl(a) for a in _split_args(args)]
        model["fns"][name] = (None if ret == "void" else c_type_to_rust(ret), params)
    return model


def parse_rust(path: str = FFI_RS) -> dict:
    src = re.sub(r"//[^\n]*", " ", open(path).read())
    model = {"consts": {}, "structs": {}, "opaque": [], "callbacks": {}, "fns": {}}
    for m in re.finditer(r"pub const (\w+): \w+ = ([^;]+);", src):
        model["consts"][m.group(1)] = int(m.group(2).replace("_", ""), 0)
    for m in re.finditer(r"#\[repr\(C\)\][^{};]*?pub struct (\w+)\s*\{(.*?)\}", src, flags=re.S):
        fields = [(f.group(1), " ".join(f.group(2).split())) for f in re.finditer(r"pub (\w+):\s*([^,]+),", m.group(2))]
        if fields == [("_opaque", "[u8; 0]")]:
            model["opaque"].append(m.group(1))
        else:
            model["structs"][m.group(1)] = fields

    def params(text):
        out = []
        for a in _split_args(text):
            if a:
                n, t = a.split(":", 1)
                out.append((" ".join(t.split()), n.strip()))
        return out

    for m in re.finditer(r"pub type (\w+) =\s*Option<unsafe extern \"C\" fn\((.*?)\)\s*->\s*(\w+)>;", src, flags=re.S):
        model["callbacks"][m.group(1)] = (m.group(3), params(m.group(2)))
    ext = re.search(r'extern "C" \{(.*)\}', src, flags=re.S)
    for m in re.finditer(r"pub fn (\w+)\((.*?)\)\s*(?:->\s*([^;]+?))?;", ext.group(1) if ext else "", flags=re.S):
        model["fns"][m.group(1)] = (" ".join(m.group(3).split()) if m.group(3) else None, params(m.group(2)))
    return model


SIZES = {"u8": (1, 1), "c_char": (1, 1), "u32": (4, 4), "i32": (4, 4), "c_int": (4, 4), "c_uint": (4, 4), "u64": (8, 8),
         "usize": (8, 8), "f64": (8, 8)}


def layout(fields, structs):
    """repr(C) layout of a parsed struct: ([(name, offset, size)], size, align)."""
    def size_align(t):
        if t.startswith("*") or t.startswith("Option<"):
            return 8, 8
        m = re.match(r"^\[(.+); (\d+)\]$", t)
        if m:
            s, a = size_align(m.group(1))
            return s * int(m.group(2)), a
        if t in SIZES:
            return SIZES[t]
        _, s, a = layout(structs[t], structs)
        return s, a

    off, align, out = 0, 1, []
    for name, t in fields:
        s, a = size_align(t)
        off = (off + a - 1) // a * a
        out.append((name, off, s))
        off += s
        align = max(align, a)
    return out, (off + align - 1) // align * align, align
