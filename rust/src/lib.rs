//! `ode_solvers`' explicit Runge-Kutta surface (Rk4, Dopri5, Dop853) over the B200 batched
//! CUDA step loop. NOT COMPILED IN THE BUILD IMAGE (no Rust toolchain there).
//!
//! Differences from the reference crate, all forced by running the RHS on a GPU:
//! * `System` is not a Rust closure: it names a device functor (`DeviceSystem::builtin`) or
//!   carries CUDA source for `system` / `solout` (`DeviceSystem::from_source`, NVRTC); the
//!   implementing struct's fields become the parameter vector.
//! * state vectors are `nalgebra::SVector<T, D>` (compile-time dimension: `Dopri5`, `Dop853`, `Rk4` for
//!   f64, `Dopri5F32`, `Dop853F32`, `Rk4F32` for f32) or `nalgebra::DVector<f64>` (`Dopri5Dyn`, `Dop853Dyn`,
//!   `Rk4Dyn`: the `*_dvector.rs` examples; the device kernel is specialised for the length it is given).
//! * `integrate_batch` is new: many trajectories, one configuration.
//!
//! `ffi.rs` is generated from `include/ode_b200.h` (tools/gen_rust_ffi.py) and checked against it by
//! tests/test_rust_ffi.py, which also checks that every `ffi::` call below passes the number of
//! arguments the ABI takes.
pub mod ffi;

pub use nalgebra::{DVector, SVector, Vector1, Vector2, Vector3, Vector4, Vector5, Vector6};
use serde::{Deserialize, Serialize};
use std::ffi::{CStr, CString};
use thiserror::Error;

/// src/dop_shared.rs:112-117
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum OutputType {
    Dense = 0,
    Sparse = 1,
    Continuous = 2,
}

/// src/dop_shared.rs:131-136
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct Stats {
    pub num_eval: u32,
    pub accepted_steps: u32,
    pub rejected_steps: u32,
}

impl std::fmt::Display for Stats {
    fn fmt(&self, f: &mut std::fmt::Formatter) -> std::fmt::Result {
        writeln!(f, "Number of function evaluations: {}", self.num_eval)?;
        writeln!(f, "Number of accepted steps: {}", self.accepted_steps)?;
        write!(f, "Number of rejected steps: {}", self.rejected_steps)
    }
}

/// src/dop_shared.rs:120-128 plus the two conditions only the batched ABI can report.
#[derive(Debug, Error)]
pub enum IntegrationError {
    #[error("Stopped at x = {x}. Need more than {n_step} steps.")]
    MaxNumStepReached { x: f64, n_step: u32 },
    #[error("Stopped at x = {x}. Step size underflow.")]
    StepSizeUnderflow { x: f64 },
    #[error("The problem seems to become stiff at x = {x}.")]
    StiffnessDetected { x: f64 },
    #[error("input outside the domain the reference handles (it panics or never returns)")]
    UnsupportedDomain,
    #[error("ode_b200: {0}")]
    Backend(String),
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::ode_b200_last_error()).to_string_lossy().into_owned() }
}

/// The device-side counterpart of `trait System<T, V>` (src/dop_shared.rs:31-42).
#[derive(Clone, Debug)]
pub struct DeviceSystem {
    sys_id: i32,
    pub dim: usize,
    pub params: Vec<f64>,
}

impl DeviceSystem {
    pub fn builtin(name: &str, params: &[f64]) -> Result<Self, IntegrationError> {
        let c = CString::new(name).unwrap();
        let id = unsafe { ffi::ode_b200_system_builtin(c.as_ptr()) };
        Self::finish(id, params)
    }

    /// `rhs_body`: body of `void rhs(double x, const double* y, double* dy, const double* p)`;
    /// `solout_body`: body of `bool solout(double x, const double* y, const double* dy, const double* p)`.
    pub fn from_source(name: &str, dim: usize, rhs_body: &str, solout_body: Option<&str>, params: &[f64])
        -> Result<Self, IntegrationError> {
        let (n, r) = (CString::new(name).unwrap(), CString::new(rhs_body).unwrap());
        let s = solout_body.map(|s| CString::new(s).unwrap());
        let id = unsafe {
            ffi::ode_b200_system_from_source(n.as_ptr(), dim as i32, params.len() as i32, r.as_ptr(),
                                             s.as_ref().map_or(std::ptr::null(), |s| s.as_ptr()))
        };
        Self::finish(id, params)
    }

    fn finish(id: i32, params: &[f64]) -> Result<Self, IntegrationError> {
        if id < 0 {
            return Err(IntegrationError::Backend(last_error()));
        }
        let dim = unsafe { ffi::ode_b200_system_dim(id) } as usize;
        Ok(Self { sys_id: id, dim, params: params.to_vec() })
    }
}

/// SolverResult (src/dop_shared.rs:46,79-109)
#[derive(Clone, Debug)]
pub struct SolverResult<T, V>(pub Vec<T>, pub Vec<V>);

impl<T, V> Default for SolverResult<T, V> {
    fn default() -> Self {
        Self(Vec::new(), Vec::new())
    }
}

impl<T, V> SolverResult<T, V> {
    pub fn new(x: Vec<T>, y: Vec<V>) -> Self {
        SolverResult(x, y)
    }
    pub fn with_capacity(n: usize) -> Self {
        SolverResult(Vec::with_capacity(n), Vec::with_capacity(n))
    }
    pub fn push(&mut self, x: T, y: V) {
        self.0.push(x);
        self.1.push(y);
    }
    pub fn get(&self) -> (&Vec<T>, &Vec<V>) {
        (&self.0, &self.1)
    }
    pub fn append(&mut self, mut other: SolverResult<T, V>) {
        self.0.append(&mut other.0);
        self.1.append(&mut other.1);
    }
}

/// src/continuous_output_model.rs:6-12,107-111: same derives, same field names, so a model serialised by the
/// reference crate deserialises here and vice versa (nalgebra's `serde-serialize` feature covers `V`).
#[derive(Clone, Debug, Default, Serialize, Deserialize)]
pub struct ContinuousOutputModel<T, V> {
    lower_bound: T,
    breakpoints: Vec<T>,
    intervals: Vec<Interval<T, V>>,
}

#[derive(Clone, Debug, Serialize, Deserialize)]
struct Interval<T, V> {
    coefficients: Vec<V>,
    step_size: T,
}

macro_rules! com_impl {
    ($t:ty) => {
        impl<V> ContinuousOutputModel<$t, V>
        where
            V: Clone + std::ops::Add<V, Output = V> + std::ops::Mul<$t, Output = V>,
        {
            pub fn new() -> Self {
                Self { lower_bound: 0.0, breakpoints: Vec::new(), intervals: Vec::new() }
            }
            pub fn bounds(&self) -> ($t, $t) {
                (self.lower_bound, self.breakpoints[self.breakpoints.len() - 1])
            }
            /// src/continuous_output_model.rs:31-56,86-104 on the host (`ode_b200_com_evaluate` is the batched
            /// device form)
            pub fn evaluate(&self, x: $t) -> Option<V> {
                if x < self.lower_bound {
                    return None;
                }
                let index = match self.breakpoints.binary_search_by(|p| p.partial_cmp(&x).unwrap()) {
                    Ok(i) => i,
                    Err(i) if i < self.intervals.len() => i,
                    _ => return None,
                };
                let lower = if index == 0 { self.lower_bound } else { self.breakpoints[index - 1] };
                let iv = &self.intervals[index];
                let theta = (x - lower) / iv.step_size;
                let theta1 = 1.0 - theta;
                let c = &iv.coefficients;
                let mut result = c[c.len() - 1].clone();
                for i in (0..c.len() - 1).rev() {
                    result = c[i].clone() + result * if i % 2 == 0 { theta } else { theta1 };
                }
                Some(result)
            }
        }
    };
}
com_impl!(f64);
com_impl!(f32);

/// Results of a batched call (SoA layouts of include/ode_b200.h).
#[derive(Clone, Debug, Default)]
pub struct BatchResult {
    pub y_final: Vec<f64>, // [dim][n]
    pub x_final: Vec<f64>,
    pub h_next: Vec<f64>,
    pub num_eval: Vec<u32>,
    pub accepted: Vec<u32>,
    pub rejected: Vec<u32>,
    pub n_step: Vec<u32>,
    pub status: Vec<i32>,
}

/// One GPU context (stream + workspace).
pub struct Context(*mut ffi::ode_b200_ctx);
impl Context {
    pub fn new(device: i32) -> Result<Self, IntegrationError> {
        let p = unsafe { ffi::ode_b200_create(device) };
        if p.is_null() { Err(IntegrationError::Backend(last_error())) } else { Ok(Self(p)) }
    }
    /// Host-buffer entry: store per-trajectory results straight into pinned host arrays (default) or
    /// copy them back by DMA (`ode_b200_set_zero_copy`).
    pub fn set_zero_copy(&self, enabled: bool) {
        unsafe { ffi::ode_b200_set_zero_copy(self.0, enabled as std::os::raw::c_int) };
    }
    /// Persistent lanes with a work queue (default) or one thread per trajectory (`ode_b200_set_work_queue`).
    pub fn set_work_queue(&self, enabled: bool) {
        unsafe { ffi::ode_b200_set_work_queue(self.0, enabled as std::os::raw::c_int) };
    }
    /// Host-buffer entry: chunked upload / integrate / copy-back pipeline; `chunk_log2` = 16 by default, 0 = off
    /// (`ode_b200_set_pipeline`).
    pub fn set_pipeline(&self, chunk_log2: i32) {
        unsafe { ffi::ode_b200_set_pipeline(self.0, chunk_log2) };
    }
    /// Large states: three lanes per trajectory (default) or one thread (`ode_b200_set_lane_split`).
    pub fn set_lane_split(&self, enabled: bool) {
        unsafe { ffi::ode_b200_set_lane_split(self.0, enabled as std::os::raw::c_int) };
    }
    /// Device time of the last step-loop kernel on the context's stream, ms.
    pub fn last_kernel_ms(&self) -> f64 {
        unsafe { ffi::ode_b200_last_kernel_ms(self.0) }
    }
    pub fn launch_count(&self) -> u64 {
        unsafe { ffi::ode_b200_launch_count(self.0) }
    }
    /// The context's CUDA stream (`cudaStream_t`) for `integrate_batch_device`.
    pub fn stream(&self) -> *mut std::os::raw::c_void {
        unsafe { ffi::ode_b200_stream(self.0) }
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::ode_b200_destroy(self.0) }
    }
}

pub fn abi_version() -> i32 {
    unsafe { ffi::ode_b200_abi_version() }
}
pub fn device_count() -> i32 {
    unsafe { ffi::ode_b200_device_count() }
}
/// (FP64 pipe thread-instructions per second, implied SM clock in MHz) of `device`.
pub fn measure_fp64_peak(device: i32) -> Result<(f64, f64), IntegrationError> {
    let (mut a, mut b) = (0.0f64, 0.0f64);
    let rc = unsafe { ffi::ode_b200_measure_fp64_peak(device, &mut a, &mut b) };
    if rc != 0 { Err(IntegrationError::Backend(last_error())) } else { Ok((a, b)) }
}
/// The library's own copy of a tableau coefficient (0-based), e.g. `tableau("a8", 11, 3)`.
pub fn tableau(which: &str, i: i32, j: i32) -> f64 {
    let c = CString::new(which).unwrap();
    unsafe { ffi::ode_b200_tableau(c.as_ptr(), i, j) }
}

impl DeviceSystem {
    pub fn n_params(&self) -> usize {
        (unsafe { ffi::ode_b200_system_n_params(self.sys_id) }) as usize
    }
    pub fn name(&self) -> String {
        unsafe { CStr::from_ptr(ffi::ode_b200_system_name(self.sys_id)).to_string_lossy().into_owned() }
    }
    /// Compile the kernel of (method, out_type) now instead of at the first launch (NVRTC; no GPU needed).
    pub fn compile(&self, method: i32, out_type: OutputType) -> Result<(), IntegrationError> {
        let rc = unsafe { ffi::ode_b200_system_compile(self.sys_id, method, out_type as i32) };
        if rc != 0 { Err(IntegrationError::Backend(last_error())) } else { Ok(()) }
    }
}

/// The batched entry point: `y0` is SoA `[dim][n]`; `params_per_traj` is `[n_params][n]` or None
/// (then `system.params` is shared by all trajectories). `n_gpus > 1` shards contiguous index ranges over
/// the GPUs of the box (`ode_b200_integrate_batch_multi`; `ctx` is then unused).
pub fn integrate_batch(ctx: &Context, system: &DeviceSystem, cfg: &ffi::ode_b200_config, n: usize, y0: &[f64],
                       params_per_traj: Option<&[f64]>, n_gpus: i32) -> Result<BatchResult, IntegrationError> {
    assert_eq!(y0.len(), system.dim * n);
    let mut r = BatchResult {
        y_final: vec![0.0; system.dim * n], x_final: vec![0.0; n], h_next: vec![0.0; n], num_eval: vec![0; n],
        accepted: vec![0; n], rejected: vec![0; n], n_step: vec![0; n], status: vec![-1; n],
    };
    let out = ffi::ode_b200_outputs {
        y_final: r.y_final.as_mut_ptr(), x_final: r.x_final.as_mut_ptr(), h_next: r.h_next.as_mut_ptr(),
        num_eval: r.num_eval.as_mut_ptr(), accepted: r.accepted.as_mut_ptr(), rejected: r.rejected.as_mut_ptr(),
        n_step: r.n_step.as_mut_ptr(), status: r.status.as_mut_ptr(), out_cap: 0, out_x: std::ptr::null_mut(),
        out_y: std::ptr::null_mut(), out_count: std::ptr::null_mut(), com_cap: 0, com_lower: std::ptr::null_mut(),
        com_break: std::ptr::null_mut(), com_step: std::ptr::null_mut(), com_coef: std::ptr::null_mut(),
        com_count: std::ptr::null_mut(),
    };
    let (pp, per) = match params_per_traj {
        Some(p) => (p.as_ptr(), 1),
        None => (system.params.as_ptr(), 0),
    };
    let rc = if n_gpus > 1 {
        unsafe { ffi::ode_b200_integrate_batch_multi(system.sys_id, cfg, n as i64, y0.as_ptr(), pp, per, n_gpus, &out) }
    } else {
        unsafe { ffi::ode_b200_integrate_batch(ctx.0, system.sys_id, cfg, n as i64, y0.as_ptr(), pp, per, &out) }
    };
    if rc != 0 { Err(IntegrationError::Backend(last_error())) } else { Ok(r) }
}

/// Device-pointer form of the batched entry: every pointer (y0, params, all of `out`) is device memory on the
/// context's GPU; the kernel is enqueued on `stream` (a `cudaStream_t`; null = the default stream) without
/// synchronising.
///
/// # Safety
/// The pointers must be valid device allocations of the documented sizes until the stream has finished.
pub unsafe fn integrate_batch_device(ctx: &Context, system: &DeviceSystem, cfg: &ffi::ode_b200_config, n: usize,
                                     y0: *const f64, params: *const f64, params_per_traj: bool,
                                     out: &ffi::ode_b200_outputs, stream: *mut std::os::raw::c_void)
                                     -> Result<(), IntegrationError> {
    let rc = ffi::ode_b200_integrate_batch_device(ctx.0, system.sys_id, cfg, n as i64, y0, params,
                                                  params_per_traj as std::os::raw::c_int, out, stream);
    if rc != 0 { Err(IntegrationError::Backend(last_error())) } else { Ok(()) }
}

/// Raw result of a single-trajectory run: samples, flattened states / coefficients, Stats, error payloads.
struct Raw<T> {
    xs: Vec<T>,
    ys: Vec<T>,
    stats: Stats,
    status: i32,
    x_final: T,
    n_step: u32,
    com_lower: T,
    com_break: Vec<T>,
    com_step: Vec<T>,
    com_coef: Vec<T>,
}

fn status_to_result(status: i32, x: f64, n_step: u32, stats: Stats) -> Result<Stats, IntegrationError> {
    match status {
        ffi::ODE_B200_MAX_NUM_STEP_REACHED => Err(IntegrationError::MaxNumStepReached { x, n_step }),
        ffi::ODE_B200_STEP_SIZE_UNDERFLOW => Err(IntegrationError::StepSizeUnderflow { x }),
        ffi::ODE_B200_STIFFNESS_DETECTED => Err(IntegrationError::StiffnessDetected { x }),
        ffi::ODE_B200_UNSUPPORTED_DOMAIN => Err(IntegrationError::UnsupportedDomain),
        _ => Ok(stats),
    }
}

/// One trajectory through the f64 host-buffer entry, growing the sample capacity until everything fits.
fn run_single(ctx: &Context, f: &DeviceSystem, cfg: &ffi::ode_b200_config, y: &[f64], m: usize, want_com: bool)
              -> Result<Raw<f64>, IntegrationError> {
    let d = y.len();
    let mut cap: usize = if cfg.out_type == ffi::ODE_B200_OUT_DENSE && cfg.dx > 0.0 {
        ((cfg.x_end - cfg.x0) / cfg.dx).max(0.0) as usize + 3
    } else { 4096 };
    loop {
        let (mut ox, mut oy) = (vec![0.0f64; cap], vec![0.0f64; cap * d]);
        let ccap = if want_com { cap } else { 0 };
        let (mut cb, mut cs, mut cc) = (vec![0.0f64; ccap], vec![0.0f64; ccap], vec![0.0f64; ccap * m * d]);
        let (mut yf, mut xf, mut hn) = (vec![0.0f64; d], [0.0f64], [0.0f64]);
        let (mut ne, mut ac, mut rj, mut ns, mut st) = ([0u32], [0u32], [0u32], [0u32], [-1i32]);
        let (mut oc, mut cn, mut cl) = ([0u32], [0u32], [0.0f64]);
        let out = ffi::ode_b200_outputs {
            y_final: yf.as_mut_ptr(), x_final: xf.as_mut_ptr(), h_next: hn.as_mut_ptr(),
            num_eval: ne.as_mut_ptr(), accepted: ac.as_mut_ptr(), rejected: rj.as_mut_ptr(),
            n_step: ns.as_mut_ptr(), status: st.as_mut_ptr(), out_cap: cap as i64,
            out_x: ox.as_mut_ptr(), out_y: oy.as_mut_ptr(), out_count: oc.as_mut_ptr(),
            com_cap: ccap as i64, com_lower: cl.as_mut_ptr(),
            com_break: cb.as_mut_ptr(), com_step: cs.as_mut_ptr(), com_coef: cc.as_mut_ptr(),
            com_count: cn.as_mut_ptr(),
        };
        let rc = unsafe {
            ffi::ode_b200_integrate_batch(ctx.0, f.sys_id, cfg, 1, y.as_ptr(), f.params.as_ptr(), 0, &out)
        };
        if rc != 0 { return Err(IntegrationError::Backend(last_error())); }
        if oc[0] as usize > cap || (want_com && cn[0] as usize > cap) { cap = (oc[0].max(cn[0])) as usize; continue; }
        let (n, q) = (oc[0] as usize, if want_com { cn[0] as usize } else { 0 });
        ox.truncate(n);
        oy.truncate(n * d);
        cb.truncate(q);
        cs.truncate(q);
        cc.truncate(q * m * d);
        return Ok(Raw {
            xs: ox, ys: oy, stats: Stats { num_eval: ne[0], accepted_steps: ac[0], rejected_steps: rj[0] },
            status: st[0], x_final: xf[0], n_step: ns[0], com_lower: cl[0], com_break: cb, com_step: cs, com_coef: cc,
        });
    }
}

/// The same through the f32 entry (`ode_b200_integrate_batch_f32`).
fn run_single_f32(ctx: &Context, f: &DeviceSystem, cfg: &ffi::ode_b200_config_f32, y: &[f32], want_com: bool)
                  -> Result<Raw<f32>, IntegrationError> {
    let (d, m) = (y.len(), 5usize);
    let params: Vec<f32> = f.params.iter().map(|&p| p as f32).collect();
    let mut cap: usize = if cfg.out_type == ffi::ODE_B200_OUT_DENSE && cfg.dx > 0.0 {
        ((cfg.x_end - cfg.x0) / cfg.dx).max(0.0) as usize + 3
    } else { 4096 };
    loop {
        let (mut ox, mut oy) = (vec![0.0f32; cap], vec![0.0f32; cap * d]);
        let ccap = if want_com { cap } else { 0 };
        let (mut cb, mut cs, mut cc) = (vec![0.0f32; ccap], vec![0.0f32; ccap], vec![0.0f32; ccap * m * d]);
        let (mut yf, mut xf, mut hn) = (vec![0.0f32; d], [0.0f32], [0.0f32]);
        let (mut ne, mut ac, mut rj, mut ns, mut st) = ([0u32], [0u32], [0u32], [0u32], [-1i32]);
        let (mut oc, mut cn, mut cl) = ([0u32], [0u32], [0.0f32]);
        let out = ffi::ode_b200_outputs_f32 {
            y_final: yf.as_mut_ptr(), x_final: xf.as_mut_ptr(), h_next: hn.as_mut_ptr(),
            num_eval: ne.as_mut_ptr(), accepted: ac.as_mut_ptr(), rejected: rj.as_mut_ptr(),
            n_step: ns.as_mut_ptr(), status: st.as_mut_ptr(), out_cap: cap as i64,
            out_x: ox.as_mut_ptr(), out_y: oy.as_mut_ptr(), out_count: oc.as_mut_ptr(),
            com_cap: ccap as i64, com_lower: cl.as_mut_ptr(),
            com_break: cb.as_mut_ptr(), com_step: cs.as_mut_ptr(), com_coef: cc.as_mut_ptr(),
            com_count: cn.as_mut_ptr(),
        };
        let rc = unsafe {
            ffi::ode_b200_integrate_batch_f32(ctx.0, f.sys_id, cfg, 1, y.as_ptr(), params.as_ptr(), 0, &out)
        };
        if rc != 0 { return Err(IntegrationError::Backend(last_error())); }
        if oc[0] as usize > cap || (want_com && cn[0] as usize > cap) { cap = (oc[0].max(cn[0])) as usize; continue; }
        let (n, q) = (oc[0] as usize, if want_com { cn[0] as usize } else { 0 });
        ox.truncate(n);
        oy.truncate(n * d);
        cb.truncate(q);
        cs.truncate(q);
        cc.truncate(q * m * d);
        return Ok(Raw {
            xs: ox, ys: oy, stats: Stats { num_eval: ne[0], accepted_steps: ac[0], rejected_steps: rj[0] },
            status: st[0], x_final: xf[0], n_step: ns[0], com_lower: cl[0], com_break: cb, com_step: cs, com_coef: cc,
        });
    }
}

/// Conversions between a state vector type and a flat slice: `SVector<T, D>` and `DVector<T>`.
pub trait StateVector<T>: Clone {
    fn as_flat(&self) -> &[T];
    fn from_flat(v: &[T]) -> Self;
}
impl<const D: usize> StateVector<f64> for SVector<f64, D> {
    fn as_flat(&self) -> &[f64] { self.as_slice() }
    fn from_flat(v: &[f64]) -> Self { SVector::<f64, D>::from_column_slice(v) }
}
impl<const D: usize> StateVector<f32> for SVector<f32, D> {
    fn as_flat(&self) -> &[f32] { self.as_slice() }
    fn from_flat(v: &[f32]) -> Self { SVector::<f32, D>::from_column_slice(v) }
}
impl StateVector<f64> for DVector<f64> {
    fn as_flat(&self) -> &[f64] { self.as_slice() }
    fn from_flat(v: &[f64]) -> Self { DVector::<f64>::from_column_slice(v) }
}

fn unpack<T: Copy, V: StateVector<T>>(raw: &Raw<T>, d: usize) -> SolverResult<T, V> {
    SolverResult(raw.xs.clone(), (0..raw.xs.len()).map(|s| V::from_flat(&raw.ys[s * d..(s + 1) * d])).collect())
}

fn fill_model<T: Copy, V: StateVector<T>>(model: &mut ContinuousOutputModel<T, V>, raw: &Raw<T>, d: usize, m: usize) {
    model.lower_bound = raw.com_lower; // set_lower_bound, dopri5.rs:279-281
    for q in 0..raw.com_break.len() {
        model.breakpoints.push(raw.com_break[q]);
        model.intervals.push(Interval {
            coefficients: (0..m).map(|r| V::from_flat(&raw.com_coef[(q * m + r) * d..(q * m + r + 1) * d])).collect(),
            step_size: raw.com_step[q],
        });
    }
}

macro_rules! stepper {
    ($name:ident, $method:expr, $m:expr) => {
        /// Single-trajectory stepper with the reference's constructors and getters (an N = 1 batch); `V` is
        /// `SVector<f64, D>` or `DVector<f64>`.
        pub struct $name<V: StateVector<f64>> {
            f: DeviceSystem,
            cfg: ffi::ode_b200_config,
            y: V,
            results: SolverResult<f64, V>,
            stats: Stats,
        }

        impl<V: StateVector<f64>> $name<V> {
            /// `new(f, x, x_end, dx, y, rtol, atol)` (dopri5.rs:76 / dop853.rs:76)
            pub fn new(f: DeviceSystem, x: f64, x_end: f64, dx: f64, y: V, rtol: f64, atol: f64) -> Self {
                let mut cfg = ffi::ode_b200_config::default();
                unsafe { ffi::ode_b200_config_new($method, x, x_end, dx, rtol, atol, &mut cfg) };
                Self { f, cfg, y, results: SolverResult::default(), stats: Stats::default() }
            }

            /// `from_param(...)` (dopri5.rs:129-146 / dop853.rs:133-150)
            #[allow(clippy::too_many_arguments)]
            pub fn from_param(f: DeviceSystem, x: f64, x_end: f64, dx: f64, y: V, rtol: f64, atol: f64,
                              safety_factor: f64, beta: f64, fac_min: f64, fac_max: f64, h_max: f64, h: f64,
                              n_max: u32, n_stiff: u32, out_type: OutputType) -> Self {
                let mut cfg = ffi::ode_b200_config::default();
                unsafe {
                    ffi::ode_b200_config_from_param($method, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min,
                                                    fac_max, h_max, h, n_max, n_stiff, out_type as i32, &mut cfg)
                };
                Self { f, cfg, y, results: SolverResult::default(), stats: Stats::default() }
            }

            fn run(&mut self, ctx: &Context, com: Option<&mut ContinuousOutputModel<f64, V>>)
                -> Result<Stats, IntegrationError> {
                let d = self.y.as_flat().len();
                let raw = run_single(ctx, &self.f, &self.cfg, self.y.as_flat(), $m, com.is_some())?;
                self.results = unpack(&raw, d);
                self.stats = raw.stats;
                if let Some(model) = com {
                    fill_model(model, &raw, d, $m);
                }
                status_to_result(raw.status, raw.x_final, raw.n_step, raw.stats)
            }

            pub fn x_out(&self) -> &Vec<f64> { self.results.get().0 }
            pub fn y_out(&self) -> &Vec<V> { self.results.get().1 }
            pub fn results(&self) -> &SolverResult<f64, V> { &self.results }
        }

        impl<V: StateVector<f64>> From<$name<V>> for SolverResult<f64, V> {
            fn from(val: $name<V>) -> Self { val.results }
        }
    };
}

stepper!(Dopri5, ffi::ODE_B200_DOPRI5, 5);
stepper!(Dop853, ffi::ODE_B200_DOP853, 8);

impl<V: StateVector<f64>> Dopri5<V> {
    /// src/dopri5.rs:255-260 (panics for OutputType::Continuous like the reference)
    pub fn integrate(&mut self, ctx: &Context) -> Result<Stats, IntegrationError> {
        if self.cfg.out_type == ffi::ODE_B200_OUT_CONTINUOUS {
            panic!("Please use `integrate_with_continuous_output_model` to compute a continuous output model.");
        }
        self.run(ctx, None)
    }
    /// src/dopri5.rs:246-252
    pub fn integrate_with_continuous_output_model(&mut self, ctx: &Context, model: &mut ContinuousOutputModel<f64, V>)
        -> Result<Stats, IntegrationError> {
        self.cfg.out_type = ffi::ODE_B200_OUT_CONTINUOUS;
        self.run(ctx, Some(model))
    }
}

impl<V: StateVector<f64>> Dop853<V> {
    /// src/dop853.rs:254
    pub fn integrate(&mut self, ctx: &Context) -> Result<Stats, IntegrationError> {
        self.run(ctx, None)
    }
    /// Extension (`ODE_B200_OUT_CONTINUOUS8`): Dop853's eight dense-output vectors per accepted step.
    pub fn integrate_with_continuous_output_model(&mut self, ctx: &Context, model: &mut ContinuousOutputModel<f64, V>)
        -> Result<Stats, IntegrationError> {
        self.cfg.out_type = ffi::ODE_B200_OUT_CONTINUOUS8;
        self.run(ctx, Some(model))
    }
}

/// src/rk4.rs -- `Rk4::new(f, x, y, x_end, step_size)` (note the argument order)
pub struct Rk4<V: StateVector<f64>> {
    inner: Dop853<V>,
}
impl<V: StateVector<f64>> Rk4<V> {
    pub fn new(f: DeviceSystem, x: f64, y: V, x_end: f64, step_size: f64) -> Self {
        let mut cfg = ffi::ode_b200_config::default();
        unsafe { ffi::ode_b200_config_rk4(x, x_end, step_size, &mut cfg) };
        Self { inner: Dop853 { f, cfg, y, results: SolverResult::default(), stats: Stats::default() } }
    }
    pub fn integrate(&mut self, ctx: &Context) -> Result<Stats, IntegrationError> { self.inner.run(ctx, None) }
    pub fn x_out(&self) -> &Vec<f64> { self.inner.x_out() }
    pub fn y_out(&self) -> &Vec<V> { self.inner.y_out() }
    pub fn results(&self) -> &SolverResult<f64, V> { self.inner.results() }
}
impl<V: StateVector<f64>> From<Rk4<V>> for SolverResult<f64, V> {
    fn from(val: Rk4<V>) -> Self { val.inner.results }
}

/// The `*_dvector.rs` examples' types: the same steppers over `DVector<f64>`.
pub type Dopri5Dyn = Dopri5<DVector<f64>>;
pub type Dop853Dyn = Dop853<DVector<f64>>;
pub type Rk4Dyn = Rk4<DVector<f64>>;

// ---------------------------------------------------------------------------------- T = f32
/// The steppers with T = f32 (`impl FloatNumber for f32`, src/dop_shared.rs:74; examples/bouncing_ball.rs:15,35-36).
pub struct StepperF32<const D: usize> {
    f: DeviceSystem,
    cfg: ffi::ode_b200_config_f32,
    y: SVector<f32, D>,
    results: SolverResult<f32, SVector<f32, D>>,
    stats: Stats,
}

impl<const D: usize> StepperF32<D> {
    fn run(&mut self, ctx: &Context, com: Option<&mut ContinuousOutputModel<f32, SVector<f32, D>>>)
        -> Result<Stats, IntegrationError> {
        let raw = run_single_f32(ctx, &self.f, &self.cfg, self.y.as_slice(), com.is_some())?;
        self.results = unpack(&raw, D);
        self.stats = raw.stats;
        if let Some(model) = com {
            fill_model(model, &raw, D, 5);
        }
        status_to_result(raw.status, raw.x_final as f64, raw.n_step, raw.stats)
    }
    pub fn x_out(&self) -> &Vec<f32> { self.results.get().0 }
    pub fn y_out(&self) -> &Vec<SVector<f32, D>> { self.results.get().1 }
    pub fn results(&self) -> &SolverResult<f32, SVector<f32, D>> { &self.results }
}
impl<const D: usize> From<StepperF32<D>> for SolverResult<f32, SVector<f32, D>> {
    fn from(val: StepperF32<D>) -> Self { val.results }
}

/// `Rk4::<f32, _, _>::new(f, x, y, x_end, step_size)` (examples/bouncing_ball.rs:36)
pub struct Rk4F32<const D: usize>(pub StepperF32<D>);
impl<const D: usize> Rk4F32<D> {
    pub fn new(f: DeviceSystem, x: f32, y: SVector<f32, D>, x_end: f32, step_size: f32) -> Self {
        let mut cfg = ffi::ode_b200_config_f32::default();
        unsafe { ffi::ode_b200_config_rk4_f32(x, x_end, step_size, &mut cfg) };
        Self(StepperF32 { f, cfg, y, results: SolverResult::default(), stats: Stats::default() })
    }
    pub fn integrate(&mut self, ctx: &Context) -> Result<Stats, IntegrationError> { self.0.run(ctx, None) }
}

macro_rules! stepper_f32 {
    ($name:ident, $method:expr) => {
        pub struct $name<const D: usize>(pub StepperF32<D>);
        impl<const D: usize> $name<D> {
            /// `new(f, x, x_end, dx, y, rtol, atol)` with T = f32 (Dopri5: uround = f32::EPSILON, dopri5.rs:89)
            pub fn new(f: DeviceSystem, x: f32, x_end: f32, dx: f32, y: SVector<f32, D>, rtol: f32, atol: f32) -> Self {
                let mut cfg = ffi::ode_b200_config_f32::default();
                unsafe { ffi::ode_b200_config_new_f32($method, x, x_end, dx, rtol, atol, &mut cfg) };
                Self(StepperF32 { f, cfg, y, results: SolverResult::default(), stats: Stats::default() })
            }
            #[allow(clippy::too_many_arguments)]
            pub fn from_param(f: DeviceSystem, x: f32, x_end: f32, dx: f32, y: SVector<f32, D>, rtol: f32, atol: f32,
                              safety_factor: f32, beta: f32, fac_min: f32, fac_max: f32, h_max: f32, h: f32,
                              n_max: u32, n_stiff: u32, out_type: OutputType) -> Self {
                let mut cfg = ffi::ode_b200_config_f32::default();
                unsafe {
                    ffi::ode_b200_config_from_param_f32($method, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min,
                                                        fac_max, h_max, h, n_max, n_stiff, out_type as i32, &mut cfg)
                };
                Self(StepperF32 { f, cfg, y, results: SolverResult::default(), stats: Stats::default() })
            }
            pub fn integrate(&mut self, ctx: &Context) -> Result<Stats, IntegrationError> { self.0.run(ctx, None) }
        }
    };
}
stepper_f32!(Dopri5F32, ffi::ODE_B200_DOPRI5);
stepper_f32!(Dop853F32, ffi::ODE_B200_DOP853);

impl<const D: usize> Dopri5F32<D> {
    pub fn integrate_with_continuous_output_model(&mut self, ctx: &Context,
                                                  model: &mut ContinuousOutputModel<f32, SVector<f32, D>>)
        -> Result<Stats, IntegrationError> {
        self.0.cfg.out_type = ffi::ODE_B200_OUT_CONTINUOUS;
        self.0.run(ctx, Some(model))
    }
}

/// Batched `ContinuousOutputModel::evaluate` on the device (`ode_b200_com_evaluate`): for trajectory `i` and
/// query `q`, `y[(i * xq.len() + q) * dim ..][..dim]` and `found[i * xq.len() + q]`.
#[allow(clippy::too_many_arguments)]
pub fn com_evaluate_batch(ctx: &Context, n: usize, dim: usize, m: usize, com_cap: usize, com_lower: &[f64],
                          com_break: &[f64], com_step: &[f64], com_coef: &[f64], com_count: &[u32], xq: &[f64])
                          -> Result<(Vec<f64>, Vec<u8>), IntegrationError> {
    let mut y = vec![0.0f64; n * xq.len() * dim];
    let mut found = vec![0u8; n * xq.len()];
    let rc = unsafe {
        ffi::ode_b200_com_evaluate(ctx.0, n as i64, dim as i32, m as i32, com_cap as i64, com_lower.as_ptr(),
                                   com_break.as_ptr(), com_step.as_ptr(), com_coef.as_ptr(), com_count.as_ptr(),
                                   xq.len() as i64, xq.as_ptr(), y.as_mut_ptr(), found.as_mut_ptr())
    };
    if rc != 0 { Err(IntegrationError::Backend(last_error())) } else { Ok((y, found)) }
}

/// Device-pointer form (`ode_b200_com_evaluate_device`): the model where `integrate_batch_device` left it, `xq`, `y`
/// and `found` on the context's GPU; enqueued on `stream` without synchronising.
///
/// # Safety
/// All pointers must be valid device allocations of the documented sizes until the stream has finished.
#[allow(clippy::too_many_arguments)]
pub unsafe fn com_evaluate_device(ctx: &Context, n: usize, dim: usize, m: usize, com_cap: usize, com_lower: *const f64,
                                  com_break: *const f64, com_step: *const f64, com_coef: *const f64,
                                  com_count: *const u32, n_query: usize, xq: *const f64, y: *mut f64, found: *mut u8,
                                  stream: *mut std::os::raw::c_void) -> Result<(), IntegrationError> {
    let rc = ffi::ode_b200_com_evaluate_device(ctx.0, n as i64, dim as i32, m as i32, com_cap as i64, com_lower, com_break,
                                               com_step, com_coef, com_count, n_query as i64, xq, y, found, stream);
    if rc != 0 { Err(IntegrationError::Backend(last_error())) } else { Ok(()) }
}
