"""The host pipeline of ode_b200_integrate_batch (chunked upload behind the running kernel, per-chunk
copy-back of finished trajectories while the rest is integrated; csrc/ode_kernel_args.h): pinned result
arrays, small chunks so that a small batch has many of them, against the oracle bit for bit."""
import ctypes as C

import numpy as np
import pytest

import ode_solvers_b200 as ob
import oracle
from ode_solvers_b200 import _abi, workloads as W
from parity_util import assert_same_outputs

pytestmark = pytest.mark.gpu


def pinned_outputs(n, dim, out_cap=0, com_cap=0, m=5):
    """an _abi.HostOutputs whose per-trajectory arrays live in pinned host memory"""
    import torch
    o = _abi.HostOutputs(n, dim, out_cap, com_cap, m)
    keep = []
    for k in ("y_final", "x_final", "h_next", "num_eval", "accepted", "rejected", "n_step", "status", "out_count",
              "com_count", "com_lower"):
        a = getattr(o, k)
        t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).pin_memory()
        keep.append(t)
        setattr(o, k, t.numpy().view(a.dtype).reshape(a.shape))
    o._keep = keep
    return o


def run_pinned(system, cfg, y0, p, chunk_log2, out_cap=0, com_cap=0, queue=True, pin_inputs=True, n_gpus=1):
    import torch
    f = ob.System.builtin(system)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    n = y0.shape[1]
    if pin_inputs:
        ty = torch.from_numpy(y0.copy()).pin_memory()
        y0 = ty.numpy()
    pp, per = None, 0
    if f.n_params:
        p = np.ascontiguousarray(p, dtype=np.float64)
        per = int(p.ndim == 2)
        pp = p.ctypes.data
    m = 5 if cfg.method == ob.DOPRI5 else 8
    out = pinned_outputs(n, f.dim, out_cap, com_cap, m)
    s = out.struct()
    lib = _abi.load()
    if n_gpus == 1:
        ctx = ob.Context(0)
        ctx.set_pipeline(chunk_log2)
        ctx.set_work_queue(queue)
        rc = lib.ode_b200_integrate_batch(ctx.handle, f.sys_id, C.byref(cfg), n, y0.ctypes.data, pp, per, C.byref(s))
        ctx.close()
    else:
        rc = lib.ode_b200_integrate_batch_multi(f.sys_id, C.byref(cfg), n, y0.ctypes.data, pp, per, n_gpus, C.byref(s))
    assert rc == 0, _abi.last_error()
    return out


CASES = [
    ("lorenz", W.lorenz_ensemble, ob.DOPRI5, ob.OUT_SPARSE, 3.0, 1e-6, 0, 0),
    ("lorenz", W.lorenz_ensemble, ob.DOP853, ob.OUT_DENSE, 2.0, 1e-7, 24, 0),
    ("lorenz", W.lorenz_ensemble, ob.DOPRI5, ob.OUT_CONTINUOUS, 1.0, 1e-5, 300, 300),
    ("lorenz", W.lorenz_ensemble, ob.RK4, ob.OUT_SPARSE, 0.3, None, 304, 0),
    ("kepler", W.kepler_sweep, ob.DOP853, ob.OUT_SPARSE, None, 1e-9, 0, 0),
    ("robertson", W.robertson_sweep, ob.DOPRI5, ob.OUT_SPARSE, 0.3, None, 0, 0),       # per-trajectory parameters
    ("bouncing_ball", W.bouncing_ball_ensemble, ob.DOPRI5, ob.OUT_DENSE, 10.0, None, 1010, 0),  # ragged early stops
    ("threebody18", W.threebody18_ensemble, ob.DOP853, ob.OUT_SPARSE, 1.5, 1e-8, 0, 0),  # three lanes per trajectory
    ("threebody18", W.threebody18_ensemble, ob.DOP853, ob.OUT_CONTINUOUS8, 1.0, 1e-8, 60, 60),
    ("threebody18", W.threebody18_ensemble, ob.DOPRI5, ob.OUT_SPARSE, 1.5, 1e-7, 0, 0),  # Dopri5 on three lanes
]


@pytest.mark.parametrize("system,gen,method,out_type,x_end,tol,cap,com", CASES)
def test_pipelined_host_entry_matches_oracle(system, gen, method, out_type, x_end, tol, cap, com):
    n = 5 * 256 + 77  # six chunks of 2^8, the last one partial
    y0, p = gen(n, offset=31)
    if system == "kepler":
        x_end = 0.6 * W.kepler_period()
    if method == ob.RK4:
        cfg = ob.config_rk4(0.0, x_end, 1e-3)
    elif tol is None:
        cfg = ob.config_new(method, 0.0, x_end, 0.01 if system == "bouncing_ball" else 0.3, 1e-2, 1e-6, out_type=out_type)
    else:
        cfg = ob.config_new(method, 0.0, x_end, x_end / 20, tol, tol, out_type=out_type)
    o = oracle.integrate_batch(system, cfg, y0, p, out_cap=cap, com_cap=com)
    for queue in (True, False):
        g = run_pinned(system, cfg, y0, p, 8, cap, com, queue=queue)
        assert_same_outputs(g, o, "pipeline %s m%d o%d queue%d" % (system, method, out_type, queue))
    g = run_pinned(system, cfg, y0, p, 0, cap, com)  # pipeline off: zero-copy result stores into the same pinned arrays
    assert_same_outputs(g, o, "no pipeline %s" % system)
    g = run_pinned(system, cfg, y0, p, 8, cap, com, pin_inputs=False)  # pageable inputs, pinned results
    assert_same_outputs(g, o, "pipeline, pageable y0 %s" % system)


def test_pipeline_edge_sizes_and_fallbacks():
    cfg = ob.config_new(ob.DOPRI5, 0.0, 1.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE)
    for n in (1, 255, 256, 257, 512, 513, 4096 + 1):
        y0, p = W.lorenz_ensemble(n, offset=n)
        o = oracle.integrate_batch("lorenz", cfg, y0, p)
        assert_same_outputs(run_pinned("lorenz", cfg, y0, p, 8), o, "pipeline n=%d" % n)
    # n_stiff = 0 (the reference panics): the fill kernel reads y0 without waiting -> the pipeline must stand down
    bad = ob.config_from_param(ob.DOPRI5, 0.0, 1.0, 0.1, 1e-6, 1e-6, 0.9, 0.04, 0.2, 10.0, 1.0, 0.0, 1000, 0, ob.OUT_SPARSE)
    y0, p = W.lorenz_ensemble(1500, offset=3)
    o = oracle.integrate_batch("lorenz", bad, y0, p)
    assert_same_outputs(run_pinned("lorenz", bad, y0, p, 8), o, "n_stiff = 0")
    assert (o.status == ob.UNSUPPORTED_DOMAIN).all()
    # repeated calls on one context reuse (stale) device buffers and flags
    ctx_runs = [run_pinned("lorenz", cfg, *W.lorenz_ensemble(2000, offset=k), 8) for k in range(3)]
    for k, g in enumerate(ctx_runs):
        assert_same_outputs(g, oracle.integrate_batch("lorenz", cfg, *W.lorenz_ensemble(2000, offset=k)), "repeat %d" % k)


def test_pipeline_at_the_default_chunk_size_large_batch():
    """2^16-trajectory chunks: 5 chunks, uniform lengths (chunks complete nearly in order) and ragged ones"""
    n = (1 << 18) + 12345
    y0, p = W.lorenz_ensemble(n, offset=7)
    cfg = ob.config_new(ob.DOPRI5, 0.0, 1.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE)
    g = run_pinned("lorenz", cfg, y0, p, 16)
    assert (g.status == ob.OK).all() and (g.x_final == 1.0).all()
    sel = np.concatenate([np.arange(0, 300), np.arange(65536 - 150, 65536 + 150), np.arange(n - 300, n), np.arange(0, n, 4099)])
    o = oracle.integrate_batch("lorenz", cfg, y0[:, sel], p)
    for k in ("accepted", "rejected", "num_eval", "n_step"):
        assert (getattr(g, k)[sel] == getattr(o, k)).all(), k
    assert (g.y_final[:, sel].view(np.uint64) == o.y_final.view(np.uint64)).all()
    assert (g.h_next[sel].view(np.uint64) == o.h_next.view(np.uint64)).all()
    y0, p = W.kepler_sweep(n, offset=1)
    cfg = ob.config_new(ob.DOP853, 0.0, 0.2 * W.kepler_period(), 60.0, 1e-8, 1e-8, out_type=ob.OUT_SPARSE)
    g = run_pinned("kepler", cfg, y0, p, 16)
    o = oracle.integrate_batch("kepler", cfg, y0[:, sel], p)
    assert (g.accepted[sel] == o.accepted).all() and (g.y_final[:, sel].view(np.uint64) == o.y_final.view(np.uint64)).all()
    assert (g.status == ob.OK).all()
