"""rust/ cannot be compiled in this image (no cargo/rustc), so its raw bindings are guarded mechanically:
the C header and rust/src/ffi.rs are parsed INDEPENDENTLY here (not with the generator's parser) and every
#[repr(C)] struct (field order, names, types), every constant and every extern "C" prototype must agree;
the committed ffi.rs must be what tools/gen_rust_ffi.py prints; struct sizes / offsets computed from the
Rust field types must equal the C compiler's (tests/c/abi_smoke.c); and every `ffi::` item lib.rs uses must
exist with the number of arguments lib.rs passes."""
import ctypes as C
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "ode_b200.h")).read()
FFI = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
LIB = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()

C2R = {"int": "c_int", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "uint8_t": "u8",
       "double": "c_double", "float": "c_float", "char": "c_char", "void": "c_void"}
RSIZE = {"c_int": 4, "i32": 4, "u32": 4, "i64": 8, "u64": 8, "u8": 1, "c_double": 8, "c_float": 4, "c_char": 1}


def c_nocomment():
    return re.sub(r"//[^\n]*", " ", re.sub(r"/\*.*?\*/", " ", HDR, flags=re.S))


def c_type_to_rust(t):
    toks = t.replace("*", " * ").split()
    const = "const" in toks
    stars = toks.count("*")
    base = [x for x in toks if x not in ("const", "*")]
    assert len(base) == 1, t
    r = C2R.get(base[0], base[0])
    return ("*const " if const else "*mut ") * stars + r if stars else r


def c_structs():
    out = {}
    for m in re.finditer(r"typedef\s+struct\s+\w+\s*\{(.*?)\}\s*(\w+)\s*;", c_nocomment(), flags=re.S):
        fields = []
        for decl in filter(None, (" ".join(d.split()) for d in m.group(1).split(";"))):
            first, rest = decl.split(None, 1) if not decl.startswith("const ") else (" ".join(decl.split()[:2]), " ".join(decl.split()[2:]))
            for nm in rest.split(","):
                nm = nm.strip()
                fields.append((nm.lstrip("* ").strip(), c_type_to_rust(first + "*" * nm.count("*"))))
        out[m.group(2)] = fields
    return out


def rust_structs():
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\]\s*(?:#\[derive\([^)]*\)\]\s*)?pub struct (\w+) \{(.*?)\n\}", FFI, flags=re.S):
        fields = re.findall(r"pub (\w+): ([^,\n]+),", m.group(2))
        out[m.group(1)] = [(a, b.strip()) for a, b in fields]
    return out


def c_protos():
    body = c_nocomment()
    body = body[body.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"\b((?:const\s+)?\w+\s*\**)\s*(ode_b200_\w+)\s*\(([^()]*)\)\s*;", body):
        args = []
        raw = " ".join(m.group(3).split())
        if raw not in ("", "void"):
            for a in raw.split(","):
                a = a.strip()
                mm = re.match(r"^(.*[\s\*])(\w+)$", a)
                if a.endswith("*") or not mm or mm.group(1).strip() in ("", "const"):
                    args.append(c_type_to_rust(a))
                else:
                    args.append(c_type_to_rust(mm.group(1)))
        ret = " ".join(m.group(1).split())
        out[m.group(2)] = (None if ret == "void" else c_type_to_rust(ret), args)
    return out


def rust_protos():
    ext = FFI[FFI.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"pub fn (\w+)\(([^)]*)\)(?:\s*->\s*([^;]+))?;", ext):
        args = [a.split(":", 1)[1].strip() for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = (m.group(3).strip() if m.group(3) else None, args)
    return out


def test_ffi_rs_is_the_generators_output():
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"]) == 0


def test_structs_match_field_for_field():
    cs, rs = c_structs(), rust_structs()
    assert set(cs) == {"ode_b200_config", "ode_b200_outputs", "ode_b200_config_f32", "ode_b200_outputs_f32"}
    for name, fields in cs.items():
        assert rs[name] == fields, (name, rs[name], fields)
    assert "ode_b200_ctx" in rs and rs["ode_b200_ctx"] == []  # opaque


def test_rust_layout_equals_the_c_compilers(tmp_path):
    """size / offset of every field as Rust's #[repr(C)] lays it out (same rule as C: align to the field's own
    alignment), against offsetof() printed by the C compiler"""
    from test_native_hosts import GCC, LIBDIR
    exe = str(tmp_path / "abi_smoke")
    subprocess.run([GCC, "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                    "-L", LIBDIR, "-lode_b200", "-o", exe], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe], check=True, capture_output=True, text=True, env=env).stdout
    c_off = {}
    for line in out.splitlines():
        if line.startswith("offset "):
            _, nm, off, size = line.split()
            c_off[nm] = (int(off), int(size))
    for sname, fields in rust_structs().items():
        off = 0
        for fname, ty in fields:
            size = C.sizeof(C.c_void_p) if ty.startswith("*") else RSIZE[ty]
            off = (off + size - 1) // size * size
            assert c_off["%s.%s" % (sname, fname)] == (off, size), (sname, fname, off, size)
            off += size


def test_constants_match():
    cn = c_nocomment()
    want = {n: int(v) for n, v in re.findall(r"#define\s+(ODE_B200_[A-Z0-9_]+)\s+(-?\d+)\b", cn)}
    for body in re.findall(r"\benum\s*\{(.*?)\}\s*;", cn, flags=re.S):
        nxt = 0
        for item in filter(None, (i.strip() for i in body.split(","))):
            if "=" in item:
                item, v = [s.strip() for s in item.split("=")]
                nxt = int(v)
            want[item] = nxt
            nxt += 1
    have = {n: int(v) for n, v in re.findall(r"pub const (ODE_B200_\w+): c_int = (-?\d+);", FFI)}
    assert have == want and len(want) >= 22


def test_every_prototype_is_bound_with_identical_types():
    cp, rp = c_protos(), rust_protos()
    from ode_solvers_b200 import _abi
    assert set(cp) == set(_abi.ABI_SYMBOLS), set(cp) ^ set(_abi.ABI_SYMBOLS)
    assert set(rp) == set(cp), set(rp) ^ set(cp)
    for name in cp:
        assert rp[name] == cp[name], (name, rp[name], cp[name])


def test_lib_rs_only_uses_bound_items_with_the_right_arity():
    rp = rust_protos()
    consts = set(re.findall(r"pub const (\w+):", FFI))
    structs = set(rust_structs())
    for item in set(re.findall(r"\bffi::(\w+)", LIB)):
        assert item in rp or item in consts or item in structs, "lib.rs uses ffi::%s, which ffi.rs does not define" % item
    # calls: count top-level commas of every `ffi::ode_b200_xxx(...)` call
    for m in re.finditer(r"\bffi::(ode_b200_\w+)\s*\(", LIB):
        name, i, depth, commas, any_tok = m.group(1), m.end(), 1, 0, False
        while depth:
            ch = LIB[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                commas += 1
            elif not ch.isspace():
                any_tok = True
            i += 1
        n_args = commas + 1 if any_tok else 0
        if LIB[m.end():i - 1].rstrip().endswith(","):
            n_args -= 1
        assert n_args == len(rp[name][1]), "lib.rs calls ffi::%s with %d arguments, the ABI takes %d" % (
            name, n_args, len(rp[name][1]))
    # struct literals of the two output structs name every field exactly once
    for sname in ("ode_b200_outputs", "ode_b200_outputs_f32"):
        fields = [f for f, _ in rust_structs()[sname]]
        for m in re.finditer(r"ffi::%s\s*\{(.*?)\n\s*\};" % sname, LIB, flags=re.S):
            used = re.findall(r"(?:^|[,{\s])(\w+):(?!:)", m.group(1))
            assert sorted(used) == sorted(fields), (sname, sorted(set(fields) ^ set(used)))
