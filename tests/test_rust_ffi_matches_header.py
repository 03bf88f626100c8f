"""The Rust crate under rust/ cannot be compiled in this image (no cargo).  This test keeps its raw bindings honest the only
way available here: every `extern "C"` declaration of rust/petal-neighbors-b200/src/ffi.rs is compared, name by name and
parameter by parameter, with the prototype of include/petal_b200.h it binds, and the #[repr(C)] mirror of pn_build_opts
with the C struct (field order, names and types).  The C header itself is checked against the built library in
test_host_side.py::test_header_symbols_exported."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "petal_b200.h")
FFI_RS = os.path.join(ROOT, "rust", "petal-neighbors-b200", "src", "ffi.rs")

# C type (const dropped, whitespace normalised) -> the Rust spelling ffi.rs may use for it
C_TO_RUST = {
    "float": {"f32"}, "double": {"f64"}, "size_t": {"usize"}, "int32_t": {"i32"}, "uint32_t": {"u32"},
    "uint64_t": {"u64"}, "void": {"()"}, "char": {"c_char"},
}


def _strip_c_comments(src):
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def _c_type(decl):
    """'const float *points' -> ('float', 1, True): base type, pointer depth, constness of the pointee"""
    decl = decl.strip()
    const = bool(re.search(r"\bconst\b", decl))
    decl = re.sub(r"\bconst\b", " ", decl)
    depth = decl.count("*")
    decl = decl.replace("*", " ")
    toks = decl.split()
    # the last token is the parameter name unless the declaration is a bare type ("void")
    base = toks[0] if len(toks) >= 1 else ""
    if toks[0] in ("struct", "enum"):
        base = toks[1]
    return base, depth, const


def _rust_type(t):
    """'*const f32' -> ('f32', 1, True);  '*mut *mut pn_tree' -> ('pn_tree', 2, False)"""
    t = t.strip()
    depth, const = 0, False
    while True:
        m = re.match(r"\*(const|mut)\s+(.*)", t)
        if not m:
            break
        if depth == 0 or m.group(1) == "const":
            const = const or m.group(1) == "const"
        depth += 1
        t = m.group(2).strip()
    t = t.replace("std::ffi::c_void", "()").replace("std::os::raw::c_char", "c_char")
    return t, depth, const


def c_prototypes():
    src = _strip_c_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"\b([A-Za-z_][\w ]*?[\s\*]+)(pn_\w+)\s*\(([^;{}()]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        if "typedef" in ret:
            continue
        params = [] if args.strip() in ("", "void") else [_c_type(a) for a in args.split(",")]
        out[name] = (_c_type(ret + " x")[:2], params)
    return out


def rust_externs():
    src = re.sub(r"//.*", "", open(FFI_RS).read())
    block = re.search(r'extern\s+"C"\s*\{(.*?)\n\}', src, flags=re.S).group(1)
    out = {}
    for m in re.finditer(r"pub fn (pn_\w+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        name, args, ret = m.group(1), m.group(2), (m.group(3) or "()").strip()
        params = []
        for a in filter(None, (s.strip() for s in args.split(","))):
            params.append(_rust_type(a.split(":", 1)[1]))
        out[name] = (_rust_type(ret)[:2], params)
    return out


def _same(ctype, rtype):
    cb, cd, cc = ctype
    rb, rd, rc = rtype
    if cd != rd:
        return False
    if cd > 0 and cc != rc and cd == 1:   # constness of a single-level pointee must agree
        return False
    return rb in C_TO_RUST.get(cb, {cb})


def test_every_rust_extern_matches_its_c_prototype():
    c, r = c_prototypes(), rust_externs()
    assert len(r) >= 25, "the extern block was not parsed"
    problems = []
    for name, (rret, rparams) in r.items():
        if name not in c:
            problems.append(f"{name}: bound in ffi.rs, not declared in petal_b200.h")
            continue
        cret, cparams = c[name]
        if not _same(cret + (False,), rret + (False,)):
            problems.append(f"{name}: return type {cret} vs {rret}")
        if len(cparams) != len(rparams):
            problems.append(f"{name}: {len(cparams)} parameters in C, {len(rparams)} in Rust")
            continue
        for i, (cp, rp) in enumerate(zip(cparams, rparams)):
            if not _same(cp, rp):
                problems.append(f"{name}: parameter {i}: C {cp} vs Rust {rp}")
    assert not problems, "\n".join(problems)


def _c_struct_fields(name):
    src = _strip_c_comments(open(HEADER).read())
    body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s\s*;" % (name, name), src, flags=re.S).group(1)
    fields = []
    for decl in filter(None, (d.strip() for d in body.split(";"))):
        m = re.match(r"(\w+)\s+(.*)", decl)
        ctype = m.group(1)
        for item in m.group(2).split(","):
            fm = re.match(r"\s*(\w+)\s*(?:\[(\d+)\])?\s*$", item)
            fields.append((fm.group(1), ctype, int(fm.group(2)) if fm.group(2) else 0))
    return fields


def _rust_struct_fields(name):
    src = re.sub(r"//.*", "", open(FFI_RS).read())
    body = re.search(r"pub struct %s\s*\{(.*?)\}" % name, src, flags=re.S).group(1)
    fields = []
    for m in re.finditer(r"pub (\w+)\s*:\s*([^,\n]+)", body):
        t = m.group(2).strip()
        am = re.match(r"\[(\w+);\s*(\d+)\]", t)
        fields.append((m.group(1), am.group(1), int(am.group(2))) if am else (m.group(1), t, 0))
    return fields


def test_build_opts_mirror_has_the_c_layout():
    cf, rf = _c_struct_fields("pn_build_opts"), _rust_struct_fields("pn_build_opts")
    assert [f[0] for f in cf] == [f[0] for f in rf], "field names / order differ"
    for (n, ct, cl), (_, rt, rl) in zip(cf, rf):
        assert rt in C_TO_RUST[ct] and cl == rl, f"pn_build_opts.{n}: C {ct}[{cl}] vs Rust {rt}[{rl}]"


def test_status_and_enum_constants_agree():
    hdr = _strip_c_comments(open(HEADER).read())
    rs = open(FFI_RS).read()
    consts = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"pub const (PN_\w+): [iu]32 = (\d+);", rs))
    assert consts, "no constants parsed from ffi.rs"
    for name, val in consts.items():
        m = re.search(r"\b%s\s*=\s*(\d+)" % name, hdr)
        assert m and int(m.group(1)) == val, f"{name}: ffi.rs says {val}, the header says {m.group(1) if m else 'nothing'}"
