"""rust/src/gpu/ffi.rs against include/ibu_b200.h (the Rust crate integration).

There is no Rust toolchain in the image, so the check is mechanical: both files are parsed into one
model and every symbol, argument type, struct field, field order, constant and enum value must agree;
the struct layouts implied by the Rust field types (repr(C)) must equal what gcc lays out for the
header; and every `ffi::` call in mod.rs must name a declared function with the declared number of
arguments.  The binding a maintainer adds to the reference crate is rust/ (see rust/README.md):
src/lib.rs:178-181 (module list), src/parallel.rs:250-296 (ParallelReader), src/io/mmap.rs:286-332,
src/io/reader.rs:510-535, src/error.rs:56-128."""
import os
import re
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ffi_model as fm  # noqa: E402

ROOT = fm.ROOT


@pytest.fixture(scope="module")
def models():
    return fm.parse_header(), fm.parse_rust()


def test_every_function_is_bound_with_the_same_signature(models):
    c, r = models
    assert len(c["fns"]) >= 70
    assert sorted(c["fns"]) == sorted(r["fns"]), set(c["fns"]) ^ set(r["fns"])
    for name, (ret, params) in c["fns"].items():
        r_ret, r_params = r["fns"][name]
        assert ret == r_ret, (name, ret, r_ret)
        assert [t for t, _ in params] == [t for t, _ in r_params], (name, params, r_params)
        assert [n for _, n in params] == [n for _, n in r_params], (name, "argument names")


def test_structs_fields_order_and_opaque_handles(models):
    c, r = models
    assert sorted(c["structs"]) == sorted(r["structs"])
    for name, fields in c["structs"].items():
        assert fields == r["structs"][name], name
    assert sorted(c["opaque"]) == sorted(r["opaque"])
    assert c["callbacks"] == r["callbacks"]


def test_constants_and_enum_values(models):
    c, r = models
    assert c["consts"] == r["consts"], {k: (c["consts"].get(k), r["consts"].get(k))
                                        for k in set(c["consts"]) | set(r["consts"])
                                        if c["consts"].get(k) != r["consts"].get(k)}
    # the error codes are the IbuError variants in declaration order (src/error.rs:56-128)
    assert [c["consts"][k] for k in ("IBU_ERR_IO", "IBU_ERR_NIFFLER", "IBU_ERR_INVALID_MAGIC", "IBU_ERR_TRUNCATED_RECORD",
                                     "IBU_ERR_INVALID_VERSION", "IBU_ERR_INVALID_BARCODE_LENGTH",
                                     "IBU_ERR_INVALID_UMI_LENGTH", "IBU_ERR_INVALID_MAP_SIZE", "IBU_ERR_INVALID_INDEX",
                                     "IBU_ERR_PROCESS")] == list(range(1, 11))


def test_struct_layouts_match_what_gcc_lays_out(models, tmp_path):
    """sizeof / offsetof of every struct as compiled from the header == the repr(C) layout of the Rust fields."""
    _, r = models
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ibu_b200.h"', "int main(void) {"]
    for name, fields in r["structs"].items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for f, _ in fields:
            lines.append(f'  printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, fields in r["structs"].items():
        offs, size, _ = fm.layout(fields, r["structs"])
        assert int(got[name]) == size, name
        for f, off, _ in offs:
            assert int(got[f"{name}.{f}"]) == off, (name, f)
    # the two wire structs stay bytemuck-identical (header.rs:44-61, record.rs:58-66)
    assert got["ibu_header_t"] == "32" and got["ibu_record_t"] == "24"


def _calls(src):
    """(name, number of top-level arguments) of every ffi::name(...) call."""
    out = []
    for m in re.finditer(r"ffi::(ibu_\w+)\s*\(", src):
        depth, i, args, cur = 1, m.end(), 0, ""
        while depth and i < len(src):
            ch = src[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            if ch == "," and depth == 1:
                args += 1 if cur.strip() else 0
                cur = ""
            elif depth:
                cur += ch
            i += 1
        args += 1 if cur.strip() else 0
        out.append((m.group(1), args))
    return out


def test_mod_rs_calls_declared_functions_with_declared_arity(models):
    _, r = models
    src = re.sub(r"//[^\n]*", " ", open(os.path.join(ROOT, "rust", "src", "gpu", "mod.rs")).read())
    calls = _calls(src)
    assert len(calls) >= 25
    for name, n_args in calls:
        assert name in r["fns"], name
        assert n_args == len(r["fns"][name][1]), (name, n_args, len(r["fns"][name][1]))
    for opener, closer in ("()", "[]", "{}"):
        assert src.count(opener) == src.count(closer), opener
    # every constant / type the wrapper names exists in ffi.rs
    names = set(r["consts"]) | set(r["structs"]) | set(r["opaque"]) | set(r["callbacks"])
    for ident in set(re.findall(r"ffi::((?:IBU|ibu)_\w+)", src)) - {c for c, _ in calls}:
        assert ident in names, ident


def test_crate_skeleton_is_complete():
    for rel in ("Cargo.toml", "build.rs", "README.md", "src/gpu/ffi.rs", "src/gpu/mod.rs"):
        assert os.path.exists(os.path.join(ROOT, "rust", rel)), rel
    cargo = open(os.path.join(ROOT, "rust", "Cargo.toml")).read()
    assert re.search(r"^gpu\s*=", cargo, flags=re.M) and 'links = "ibu_b200"' in cargo
    assert "rustc-link-lib=dylib=ibu_b200" in open(os.path.join(ROOT, "rust", "build.rs")).read()


def test_ffi_rs_is_what_the_generator_renders(tmp_path):
    """ffi.rs was not edited by hand after the header changed (or the generator after ffi.rs)."""
    before = open(fm.FFI_RS).read()
    subprocess.run([sys.executable, os.path.join(ROOT, "rust", "gen_ffi.py")], check=True, capture_output=True)
    after = open(fm.FFI_RS).read()
    if after != before:
        open(fm.FFI_RS, "w").write(before)
    assert after == before, "rust/src/gpu/ffi.rs is stale: run `python rust/gen_ffi.py`"
