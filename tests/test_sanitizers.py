"""Sanitizer tier (compute-sanitizer is closed on the GPU pool, so this is the part that CAN run):
 * the CPU oracle rebuilt with -fsanitize=address,undefined and driven through every system x method x
   output mode (f64 and f32), capacity overflows and error exits -- and its results must equal the normal
   build bit for bit (-O1 vs -O2 as well);
 * the HOST side of libode_b200.so (ABI layer, registry, constructors, NVRTC source registration and
   compilation, argument validation, the no-device paths) rebuilt with the same sanitizers and driven
   without a GPU; on a GPU box the same build also runs whole integrations (tests marked gpu).
Both run in a subprocess with the sanitizer runtime preloaded."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
SAN = "-fsanitize=address,undefined"


def _runtime():
    p = subprocess.run([GCC, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    return p if os.path.isabs(p) and os.path.exists(p) else None


def _env(extra):
    rt = _runtime()
    env = dict(os.environ)
    env.update(extra)
    env["LD_PRELOAD"] = rt
    # python itself leaks by design; CUDA reserves address space ASan wants to protect
    env["ASAN_OPTIONS"] = "detect_leaks=0:protect_shadow_gap=0:abort_on_error=0:exitcode=87"
    env["UBSAN_OPTIONS"] = "print_stacktrace=1:halt_on_error=1:exitcode=88"
    return env


ORACLE_DRIVER = r"""
import hashlib, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
import oracle
from ode_solvers_b200 import workloads as W
h = hashlib.sha256()
def upd(o):
    for k in ("y_final","x_final","h_next","num_eval","accepted","rejected","n_step","status","out_count","com_count",
              "com_lower","out_x","out_y","com_break","com_step","com_coef"):
        a = getattr(o, k, None)
        if a is not None: h.update(np.ascontiguousarray(a).tobytes())
ens = {"lorenz": W.lorenz_ensemble, "kepler": W.kepler_sweep, "cr3bp": W.cr3bp_ensemble, "robertson": W.robertson_sweep,
       "bouncing_ball": W.bouncing_ball_ensemble, "threebody18": W.threebody18_ensemble}
for s, g in ens.items():
    y0, p = g(12, offset=3)
    y0[0, 5] = np.inf; y0[-1, 6] = np.nan
    xe = {"lorenz": 2.0, "kepler": 0.5 * W.kepler_period(), "cr3bp": 2.0, "robertson": 0.3, "bouncing_ball": 10.0,
          "threebody18": 1.0}[s]
    for m in (1, 2):
        for ot in (0, 1, 2, 3):
            if ot == 3 and m == 1: continue
            upd(oracle.integrate_batch(s, oracle.config_new(m, 0.0, xe, xe / 37, 1e-6, 1e-8, out_type=ot), y0, p, out_cap=60, com_cap=50))
            upd(oracle.integrate_batch(s, oracle.config_from_param(m, 0.0, xe, xe / 11, 1e-5, 1e-8, 0.85, 0.07, 0.25, 7.0, xe / 3,
                                                                   0.0, 40, 3, ot), y0, p, out_cap=7, com_cap=5))
    upd(oracle.integrate_batch(s, oracle.config_rk4(0.0, xe, xe / 100), y0, p, out_cap=90))
for s, g in (("lorenz", W.lorenz_ensemble), ("robertson", W.robertson_sweep), ("bouncing_ball", W.bouncing_ball_ensemble)):
    y0, p = g(12, offset=1)
    xe = {"lorenz": 1.0, "robertson": 0.3, "bouncing_ball": 10.0}[s]
    for m in (1, 2):
        for ot in (0, 1, 2):
            upd(oracle.integrate_batch_f32(s, oracle.config_new_f32(m, 0.0, xe, xe / 20, 1e-3, 1e-5, out_type=ot), y0, p, out_cap=30, com_cap=25))
    upd(oracle.integrate_batch_f32(s, oracle.config_rk4_f32(0.0, xe, xe / 50), y0, p, out_cap=20))
for s in ("test_rk4_1", "test_rk4_2", "test_exp", "test_cubic", "test_exp_stop", "test_cubic_stop"):
    y0 = np.linspace(0.2, 1.2, 7).reshape(1, -1)
    p = [0.5] if s.endswith("stop") else None
    for m in (1, 2):
        upd(oracle.integrate_batch(s, oracle.config_new(m, 0.0, 1.0, 0.1, 1e-8, 1e-8), y0, p, out_cap=20))
print("ORACLE_HASH", h.hexdigest())
"""


def test_oracle_under_asan_and_ubsan_is_clean_and_bit_identical(tmp_path):
    if not _runtime():
        pytest.skip("libasan.so not found next to %s" % GCC)
    lib = str(tmp_path / "libode_oracle_san.so")
    src = [os.path.join(ROOT, "oracle", f) for f in ("ode_oracle.c", "ode_oracle_systems.c")]
    subprocess.run([GCC, "-O1", "-g", "-std=c99", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", SAN,
                    "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-shared", "-o", lib] + src + ["-lm"],
                   check=True)
    code = ORACLE_DRIVER % {"root": ROOT}
    plain = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert plain.returncode == 0, plain.stderr[-2000:]
    san = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900,
                         env=_env({"ODE_ORACLE_LIB": lib, "OMP_NUM_THREADS": "2"}))
    assert san.returncode == 0, "sanitizer report:\n" + san.stderr[-4000:]
    assert "ERROR: AddressSanitizer" not in san.stderr and "runtime error" not in san.stderr, san.stderr[-4000:]
    assert san.stdout.strip().splitlines()[-1] == plain.stdout.strip().splitlines()[-1], (san.stdout, plain.stdout)


def _build_product_host_with_sanitizers():
    """libode_b200.so with its two host translation units (the ABI layer and the NVRTC driver) compiled with
    ASan + UBSan; the kernel instantiation units are reused from the normal build. Cached under variants/."""
    csrc = os.path.join(ROOT, "ode_solvers_b200", "csrc")
    out = os.path.join(ROOT, "variants", "libode_hostsan.so")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    deps.append(os.path.join(ROOT, "include", "ode_b200.h"))
    have_objects = os.path.isdir(os.path.join(csrc, "build")) and os.listdir(os.path.join(csrc, "build"))
    if os.path.exists(out) and (not have_objects or os.path.getmtime(out) >= max(os.path.getmtime(d) for d in deps)):
        return out  # (a GPU box gets the built library with the snapshot, not the object files)
    if not have_objects:
        pytest.skip("no object files of the normal build to link the sanitized host units with")
    from ode_solvers_b200 import _abi
    _abi.load()  # the normal build must exist (objects are shared)
    # nvcc splits -Xcompiler values at commas: one flag per sanitizer
    xs = "-Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined"
    flags = xs + " -Xcompiler -fno-omit-frame-pointer -Xcompiler -fno-sanitize-recover=undefined"
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "variant.sh"), "hostsan", flags, "ode_b200", "ode_jit"],
                       capture_output=True, text=True, env=dict(os.environ, VARIANT_LINK_FLAGS=xs))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return out


HOST_DRIVER = r"""
import ctypes as C, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
import ode_solvers_b200 as ob
from ode_solvers_b200 import _abi
lib = _abi.load()
assert lib.ode_b200_abi_version() == 2
for nm in ("lorenz", "kepler", "threebody18", "no_such"):
    sid = lib.ode_b200_system_builtin(nm.encode())
    if sid >= 0:
        assert lib.ode_b200_system_dim(sid) > 0 and lib.ode_b200_system_name(sid).decode() == nm
assert lib.ode_b200_system_dim(12345) < 0 and lib.ode_b200_system_name(-3) is None
for m in (ob.DOPRI5, ob.DOP853, 7):
    c = _abi.Config(); lib.ode_b200_config_new(m, 0.0, 1.0, 0.1, 1e-6, 1e-6, C.byref(c))
    c32 = _abi.ConfigF32(); lib.ode_b200_config_new_f32(m, 0.0, 1.0, 0.1, 1e-3, 1e-3, C.byref(c32))
assert lib.ode_b200_config_new(ob.DOPRI5, 0.0, 1.0, 0.1, 1e-6, 1e-6, None) < 0
# NVRTC registration + compilation of a user system (needs no GPU), incl. a source error and long names
f = ob.System.from_source("pendulum" * 20, 2, 1, "dy[0] = y[1]; dy[1] = -p[0] * sin(y[0]);", "return y[0] > 3.0;")
for m, ot in ((ob.RK4, ob.OUT_SPARSE), (ob.DOPRI5, ob.OUT_DENSE), (ob.DOP853, ob.OUT_CONTINUOUS8)):
    f.compile(m, ot)
try:
    ob.System.from_source("broken", 1, 0, "dy[0] = undefined_symbol;").compile(ob.DOPRI5, ob.OUT_SPARSE)
    raise SystemExit("a broken source compiled")
except ob.Ode200Error as e:
    assert "undefined_symbol" in str(e)
assert lib.ode_b200_system_from_source(None, 1, 0, b"x", None) < 0
assert lib.ode_b200_system_from_source(b"big", 99, 0, b"x", None) < 0
assert np.isnan(lib.ode_b200_tableau(b"a8", 99, 0)) and lib.ode_b200_tableau(b"c5", 1, 0) == 0.2
cfg = ob.config_new(ob.DOPRI5, 0.0, 1.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE)
n = 257
y0 = np.ones((3, n)); p = np.array([10.0, 8.0 / 3.0, 28.0])
out = _abi.HostOutputs(n, 3, 16, 0); s = out.struct()
if ob.device_count() == 0:
    assert lib.ode_b200_create(0) is None and "no CPU fallback" in _abi.last_error()
    assert lib.ode_b200_integrate_batch(None, 0, C.byref(cfg), n, y0.ctypes.data, p.ctypes.data, 0, C.byref(s)) < 0
    assert lib.ode_b200_integrate_batch_multi(0, C.byref(cfg), n, y0.ctypes.data, p.ctypes.data, 0, 1, C.byref(s)) == -2
    print("HOST_OK no-gpu")
else:
    import oracle
    from ode_solvers_b200 import workloads as W
    ctx = ob.Context(0)
    for sysname, gen, m, ot, cap, com in (("lorenz", W.lorenz_ensemble, ob.DOPRI5, ob.OUT_CONTINUOUS, 1001, 1001),
                                          ("lorenz", W.lorenz_ensemble, ob.RK4, ob.OUT_SPARSE, 203, 0),
                                          ("kepler", W.kepler_sweep, ob.DOP853, ob.OUT_DENSE, 37, 0),
                                          ("threebody18", W.threebody18_ensemble, ob.DOP853, ob.OUT_CONTINUOUS8, 50, 50)):
        y0, p = gen(n, offset=2)
        xe = {"lorenz": 2.0, "kepler": 2000.0, "threebody18": 1.0}[sysname]
        cfg = ob.config_rk4(0.0, 0.2, 1e-3) if m == ob.RK4 else ob.config_new(m, 0.0, xe, xe / 31, 1e-6, 1e-6, out_type=ot)
        g = ob.integrate_batch(ob.System.builtin(sysname), cfg, y0, p, out_cap=cap, com_cap=com, ctx=ctx)
        o = oracle.integrate_batch(sysname, cfg, y0, p, out_cap=cap, com_cap=com)
        assert np.array_equal(g.y_final.view(np.uint64), o.y_final.view(np.uint64)) and np.array_equal(g.accepted, o.accepted)
        for i in range(0, n, 17):
            cnt = int(min(g.out_count[i], cap))
            assert np.array_equal(g.out_y[i, :cnt].view(np.uint64), o.out_y[i, :cnt].view(np.uint64))
    g = ob.integrate_batch(f.with_params([9.81]), ob.config_new(ob.DOPRI5, 0.0, 2.0, 0.1, 1e-6, 1e-6), np.ones((2, 33)), out_cap=30)
    assert (g.status == 0).all()
    ctx.close()
    print("HOST_OK gpu")
"""


def _run_host_driver():
    lib = _build_product_host_with_sanitizers()
    r = subprocess.run([sys.executable, "-c", HOST_DRIVER % {"root": ROOT}], capture_output=True, text=True, timeout=900,
                       env=_env({"ODE_B200_LIB": lib}))
    return r


def test_product_host_layer_under_asan_and_ubsan_without_a_gpu():
    if not _runtime():
        pytest.skip("libasan.so not found")
    import ode_solvers_b200 as ob
    if ob.device_count() > 0:
        pytest.skip("a GPU is present: the gpu-marked variant runs the larger driver")
    r = _run_host_driver()
    assert r.returncode == 0, "rc %d\n%s" % (r.returncode, r.stderr[-4000:])
    assert "HOST_OK" in r.stdout and "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-4000:]


@pytest.mark.gpu
def test_product_host_layer_under_asan_and_ubsan_on_the_gpu():
    """whole integrations through the sanitized host layer: workspace growth, padded row pitches and their
    strided copy-back, zero-copy probing, the NVRTC launch path"""
    if not _runtime():
        pytest.skip("libasan.so not found")
    r = _run_host_driver()
    if r.returncode != 0 and ("cudaErrorMemoryAllocation" in r.stderr or "out of memory" in r.stderr
                              or "no CUDA device available" in r.stderr):
        pytest.skip("CUDA does not initialise under ASan's address-space layout on this box")
    assert r.returncode == 0, "rc %d\n%s" % (r.returncode, r.stderr[-4000:])
    assert "HOST_OK gpu" in r.stdout and "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr
