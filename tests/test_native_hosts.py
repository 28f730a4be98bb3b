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
    return subprocess.run([exe], check=True, capture_output=True, text=True, env=env).stdout


def test_c_header_links_and_struct_layout(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    out = _build_and_run(tmp_path, [GCC, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                                    os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-L", LIBDIR, "-lode_b200",
                                    "-o", exe], exe)
    f = out.split()
    assert f[1] == "1"
    assert int(f[3]) == C.sizeof(_abi.Config) and int(f[5]) == C.sizeof(_abi.Outputs)
    assert "alpha %.17g" % (0.2 - 0.04 * 0.75) in out and "n_max 100000" in out


def test_cpp_host_mirror_compiles_and_runs(tmp_path):
    exe = str(tmp_path / "cpp_host_smoke")
    out = _build_and_run(tmp_path, [GXX, "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-I",
                                    os.path.join(LIBDIR, "cpp"), os.path.join(ROOT, "tests", "cpp_host_smoke.cpp"),
                                    "-L", LIBDIR, "-lode_b200", "-o", exe], exe)
    assert "FAIL" not in out


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu(tmp_path):
    exe = str(tmp_path / "cpp_host_smoke")
    out = _build_and_run(tmp_path, [GXX, "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I",
                                    os.path.join(LIBDIR, "cpp"), os.path.join(ROOT, "tests", "cpp_host_smoke.cpp"),
                                    "-L", LIBDIR, "-lode_b200", "-o", exe], exe)
    assert "dopri5 test1" in out and "FAIL" not in out and out.count("ok") >= 3
