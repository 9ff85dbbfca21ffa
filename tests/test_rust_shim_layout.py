"""N3 without rustc: the Rust binding a maintainer adds (integration/rust/rtb200_sys.rs) must describe the same bytes as
include/rtb200.h. This test parses the Rust file (`#[repr(C)]` structs: field order, Rust types -> size / alignment under
the C layout rules) and the C header (typedef structs, through the ctypes mirrors that every GPU test already calls
through), and compares field names, offsets, sizes and struct alignment; then it compares the `extern "C"` signatures
with the header's prototypes. It also pins INTEGRATION.md's code blocks to the files (no second copy to drift)."""
import ctypes as C
import os
import re

import rtb200 as R

RUST_PRIMS = {"f64": (8, 8), "f32": (4, 4), "u64": (8, 8), "i64": (8, 8), "u32": (4, 4), "i32": (4, 4), "u8": (1, 1), "c_int": (4, 4), "c_char": (1, 1)}
C_TO_RUST = {"double": "f64", "float": "f32", "uint64_t": "u64", "int64_t": "i64", "uint32_t": "u32", "int32_t": "i32", "uint8_t": "u8", "int": "c_int", "char": "c_char"}


def _rust_structs(text):
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\](?:\s*#\[derive\([^)]*\)\])?\s*pub struct (\w+)\s*\{(.*?)\}", text, flags=re.S):
        fields = [(f.group(1), f.group(2).strip()) for f in re.finditer(r"pub (\w+)\s*:\s*([^,}]+)", m.group(2))]
        out[m.group(1)] = fields
    return out


def _layout(ty, structs, cache):
    """(size, align) of a Rust type under repr(C)."""
    ty = ty.strip()
    if ty in RUST_PRIMS:
        return RUST_PRIMS[ty]
    if ty.startswith("*const") or ty.startswith("*mut"):
        return (8, 8)
    m = re.fullmatch(r"\[(\w+);\s*(\d+)\]", ty)
    if m:
        s, a = _layout(m.group(1), structs, cache)
        return (s * int(m.group(2)), a)
    if ty not in cache:
        off, align, offs = 0, 1, []
        for name, fty in structs[ty]:
            s, a = _layout(fty, structs, cache)
            off = (off + a - 1) // a * a
            offs.append((name, off, s))
            off += s
            align = max(align, a)
        cache[ty] = ((off + align - 1) // align * align, align, offs)
    return cache[ty][0], cache[ty][1]


def test_rust_structs_have_the_layout_of_the_c_header(repo):
    text = open(os.path.join(repo, "integration", "rust", "rtb200_sys.rs")).read()
    structs = _rust_structs(text)
    expected = ["rt_vec3", "rt_camera", "rt_sphere", "rt_image", "rt_sky", "rt_scene", "rt_stats", "rt_options"]
    assert sorted(structs) == sorted(expected)
    cache = {}
    for name in expected:
        size, align = _layout(name, structs, cache)
        ct = getattr(R, name)
        assert size == C.sizeof(ct), (name, size, C.sizeof(ct))
        assert align == C.alignment(ct), name
        rust_fields = cache[name][2]
        c_fields = [(f[0], getattr(ct, f[0]).offset, getattr(ct, f[0]).size) for f in ct._fields_]
        assert rust_fields == c_fields, (name, rust_fields, c_fields)


def _c_structs(header):
    """Field lists of the header's typedef structs: [(c type, name, array length | None)]."""
    txt = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    out = {}
    for m in re.finditer(r"typedef struct\s*\{(.*?)\}\s*(\w+)\s*;", txt, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            tm = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.*)", decl)
            base = tm.group(2) + ("*" if tm.group(3) else "")
            for nm in tm.group(4).split(","):
                nm = nm.strip()
                ptr = nm.startswith("*")
                nm = nm.lstrip("* ")
                am = re.fullmatch(r"(\w+)\[(\d+)\]", nm)
                fields.append((base + ("*" if ptr else ""), am.group(1) if am else nm, int(am.group(2)) if am else None))
        out[m.group(2)] = fields
    return out


def test_rust_field_types_match_the_header_and_the_ctypes_mirrors_match_the_header(repo):
    header = open(os.path.join(repo, "include", "rtb200.h")).read()
    cs = _c_structs(header)
    rs = _rust_structs(open(os.path.join(repo, "integration", "rust", "rtb200_sys.rs")).read())
    for name, rfields in rs.items():
        cfields = cs[name]
        assert [f[0] for f in rfields] == [f[1] for f in cfields], name          # same names, same order
        assert [f[0] for f in getattr(R, name)._fields_] == [f[1] for f in cfields], name   # ... also in the ctypes mirror the tests call through
        for (rn, rty), (cty, cn, arr) in zip(rfields, cfields):
            if cty.endswith("*"):
                assert rty.startswith("*const"), (name, rn)
                continue
            want = C_TO_RUST.get(cty, cty)
            assert rty == (f"[{want}; {arr}]" if arr else want), (name, rn, rty, cty)


def test_extern_signatures_match_the_header(repo):
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(repo, "include", "rtb200.h")).read(), flags=re.S)
    rust = open(os.path.join(repo, "integration", "rust", "rtb200_sys.rs")).read()
    ext = re.search(r'extern "C"\s*\{(.*?)\n\}', rust, flags=re.S).group(1)
    fns = re.findall(r"pub fn (\w+)\((.*?)\)\s*->\s*([^;]+);", ext)
    assert sorted(f[0] for f in fns) == ["rtb200_last_error", "rtb200_render_rgb8", "rtb200_render_rgb8_multi"]
    for name, args, ret in fns:
        proto = re.search(r"([\w\s\*]+?)\b" + name + r"\s*\((.*?)\)\s*;", header, flags=re.S)
        c_ret = proto.group(1).strip()
        c_args = [a.strip() for a in proto.group(2).split(",") if a.strip() and a.strip() != "void"]
        r_args = [a.split(":", 1)[1].strip() for a in args.split(",") if a.strip()]
        assert len(c_args) == len(r_args), name
        for ca, ra in zip(c_args, r_args):
            cm = re.match(r"(const\s+)?(\w+)\s*(\*?)", ca)
            if cm.group(3):
                assert ra == ("*const " if cm.group(1) else "*mut ") + C_TO_RUST.get(cm.group(2), cm.group(2)), (name, ca, ra)
            else:
                assert ra == C_TO_RUST.get(cm.group(2), cm.group(2)), (name, ca, ra)
        assert ret.strip() == ("*const c_char" if "char" in c_ret else "c_int"), name


def test_integration_md_points_at_the_files_instead_of_copying_them(repo):
    md = open(os.path.join(repo, "INTEGRATION.md")).read()
    for f in ("build.rs", "rtb200_sys.rs", "render_replacement.rs"):
        assert f"integration/rust/{f}" in md
    assert "#[repr(C)]" not in md        # the struct definitions live in ONE place: integration/rust/rtb200_sys.rs
