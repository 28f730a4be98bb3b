"""The C header is C99-clean and links; its structs have the layout the ctypes mirror assumes;
the C++ host mirror (ode_solvers_b200/cpp/ode_b200.hpp) compiles against it."""
import ctypes as C
import os
import subprocess

import pytest

from ode_solvers_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "ode_solvers_b200")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build_and_run(tmp_path, cmd, exe):
    _abi.load()
    subprocess.run(cmd, check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_c_header_links_and_struct_layout(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    out = _build_and_run(tmp_path, [GCC, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                                    os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-L", LIBDIR, "-lode_b200",
                                    "-o", exe], exe)
    f = out.split()
    assert int(f[1]) == _abi.load().ode_b200_abi_version()
    assert int(f[3]) == C.sizeof(_abi.Config) and int(f[5]) == C.sizeof(_abi.Outputs)
    # every field: offset and size of the C struct == the ctypes mirror (tests/oracle.py shares these mirrors
    # with the product wrapper, so a layout bug would otherwise cancel on both sides of a parity test)
    mirrors = {"ode_b200_config": _abi.Config, "ode_b200_outputs": _abi.Outputs, "ode_b200_outputs_f32": _abi.OutputsF32,
               "ode_b200_config_f32": _abi.ConfigF32}
    seen = {k: [] for k in mirrors}
    for line in out.splitlines():
        if line.startswith("offset "):
            _, name, off, size = line.split()
            st, field = name.split(".")
            d = getattr(mirrors[st], field)
            assert (d.offset, d.size) == (int(off), int(size)), (name, d.offset, d.size, off, size)
            seen[st].append(field)
    for st, cls in mirrors.items():
        assert seen[st] == [n for n, _ in cls._fields_], (st, seen[st])
    assert "sizeof_outputs_f32 %d" % C.sizeof(_abi.OutputsF32) in out
    assert "alpha %.17g" % (0.2 - 0.04 * 0.75) in out and "n_max 100000" in out


def test_cpp_host_mirror_compiles_and_runs(tmp_path):
    exe = str(tmp_path / "cpp_host_smoke")
    out = _build_and_run(tmp_path, [GXX, "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-I",
                                    os.path.join(LIBDIR, "cpp"), os.path.join(ROOT, "tests", "cpp_host_smoke.cpp"),
                                    "-L", LIBDIR, "-lode_b200", "-o", exe], exe)
    assert "FAIL" not in out


def test_cpp_host_mirror_binds_every_export():
    """like tests/test_rust_ffi.py for the Rust crate: the C++ mirror wraps every entry point the header declares
    (the ode_b200_debug_* test hooks aside), so a new export cannot be forgotten on the compiled-host side"""
    import re
    hpp = open(os.path.join(LIBDIR, "cpp", "ode_b200.hpp")).read()
    missing = [sym for sym in _abi.ABI_SYMBOLS if "_debug_" not in sym and not re.search(r"\b%s\(" % sym, hpp)]
    assert not missing, missing


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu(tmp_path):
    exe = str(tmp_path / "cpp_host_smoke")
    out = _build_and_run(tmp_path, [GXX, "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I",
                                    os.path.join(LIBDIR, "cpp"), os.path.join(ROOT, "tests", "cpp_host_smoke.cpp"),
                                    "-L", LIBDIR, "-lode_b200", "-o", exe], exe)
    assert "dopri5 test1" in out and "FAIL" not in out and out.count(": ok") >= 15 and "0 checks failed" in out


def test_save_formats_like_rust_display(tmp_path):
    """ob.save mirrors the examples' save() (examples/lorenz.rs:56-84): `{}` of an f64 is the shortest
    round-trip digits, positional, without a trailing `.0`"""
    from ode_solvers_b200.api import _rust_display_f64 as f, save
    assert f(1.0) == "1" and f(-0.0) == "-0" and f(0.1) == "0.1" and f(1.5) == "1.5"
    assert f(1e-7) == "0.0000001" and f(1e21) == "1000000000000000000000" and f(123456789.125) == "123456789.125"
    assert f(5e-324) == "0." + "0" * 323 + "5"
    assert f(float("nan")) == "NaN" and f(float("inf")) == "inf" and f(float("-inf")) == "-inf"
    assert f(-5007.248417988539) == "-5007.248417988539" and f(1.0 / 3.0) == "0.3333333333333333"
    p = tmp_path / "outputs" / "x.dat"
    save([0.0, 0.001], [[1.0, 1.0, 1.0], [1.0000000000000002, 2.5, -3e-5]], str(p))
    assert p.read_text() == "0, 1, 1, 1\n0.001, 1.0000000000000002, 2.5, -0.00003\n"
