//! `ode_solvers`' explicit Runge-Kutta surface (Rk4, Dopri5, Dop853) over the B200 batched
//! CUDA step loop. NOT COMPILED IN THE BUILD IMAGE (no Rust toolchain there).
//!
//! Differences from the reference crate, all forced by running the RHS on a GPU:
//! * `System` is not a Rust closure: it names a device functor (`DeviceSystem::builtin`) or
//!   carries CUDA source for `system` / `solout` (`DeviceSystem::from_source`, NVRTC); the
//!   implementing struct's fields become the parameter vector.
//! * state vectors are `nalgebra::SVector<f64, D>` (f64 only, compile-time dimension).
//! * `integrate_batch` is new: many trajectories, one configuration.
pub mod ffi;

use nalgebra::SVector;
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
#[derive(Clone, Debug, Default)]
pub struct SolverResult<const D: usize>(pub Vec<f64>, pub Vec<SVector<f64, D>>);

impl<const D: usize> SolverResult<D> {
    pub fn get(&self) -> (&Vec<f64>, &Vec<SVector<f64, D>>) {
        (&self.0, &self.1)
    }
    pub fn append(&mut self, mut other: SolverResult<D>) {
        self.0.append(&mut other.0);
        self.1.append(&mut other.1);
    }
}

/// src/continuous_output_model.rs:8-12,107-111 (same field names for serde compatibility)
#[derive(Clone, Debug, Default)]
pub struct ContinuousOutputModel<const D: usize> {
    pub lower_bound: f64,
    pub breakpoints: Vec<f64>,
    pub intervals: Vec<Interval<D>>,
}

#[derive(Clone, Debug)]
pub struct Interval<const D: usize> {
    pub coefficients: Vec<SVector<f64, D>>,
    pub step_size: f64,
}

impl<const D: usize> ContinuousOutputModel<D> {
    pub fn bounds(&self) -> (f64, f64) {
        (self.lower_bound, self.breakpoints[self.breakpoints.len() - 1])
    }
    /// Host evaluation, the reference's algorithm verbatim in spirit
    /// (src/continuous_output_model.rs:31-56,86-104); `ode_b200_com_evaluate` is the batched device form.
    pub fn evaluate(&self, x: f64) -> Option<SVector<f64, D>> {
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
        let mut result = c[c.len() - 1];
        for i in (0..c.len() - 1).rev() {
            result = c[i] + result * if i % 2 == 0 { theta } else { theta1 };
        }
        Some(result)
    }
}

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
}
impl Context {
    /// Host-buffer entry: store per-trajectory results straight into pinned host arrays (default) or
    /// copy them back by DMA (`ode_b200_set_zero_copy`).
    pub fn set_zero_copy(&self, enabled: bool) {
        unsafe { ffi::ode_b200_set_zero_copy(self.0, enabled as std::os::raw::c_int) };
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::ode_b200_destroy(self.0) }
    }
}

/// The batched entry point: `y0` is SoA `[dim][n]`; `params_per_traj` is `[n_params][n]` or None
/// (then `system.params` is shared by all trajectories).
pub fn integrate_batch(ctx: &Context, system: &DeviceSystem, cfg: &ffi::ode_b200_config, n: usize, y0: &[f64],
                       params_per_traj: Option<&[f64]>) -> Result<BatchResult, IntegrationError> {
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
    let rc = unsafe { ffi::ode_b200_integrate_batch(ctx.0, system.sys_id, cfg, n as i64, y0.as_ptr(), pp, per, &out) };
    if rc != 0 { Err(IntegrationError::Backend(last_error())) } else { Ok(r) }
}

macro_rules! stepper {
    ($name:ident, $method:expr, $m:expr) => {
        /// Single-trajectory stepper with the reference's constructor and getters (an N = 1 batch).
        pub struct $name<const D: usize> {
            f: DeviceSystem,
            cfg: ffi::ode_b200_config,
            y: SVector<f64, D>,
            results: SolverResult<D>,
            stats: Stats,
        }

        impl<const D: usize> $name<D> {
            /// `new(f, x, x_end, dx, y, rtol, atol)` (dopri5.rs:76 / dop853.rs:76)
            pub fn new(f: DeviceSystem, x: f64, x_end: f64, dx: f64, y: SVector<f64, D>, rtol: f64, atol: f64) -> Self {
                let mut cfg = ffi::ode_b200_config::default();
                unsafe { ffi::ode_b200_config_new($method, x, x_end, dx, rtol, atol, &mut cfg) };
                Self { f, cfg, y, results: SolverResult::default(), stats: Stats::default() }
            }

            /// `from_param(...)` (dopri5.rs:129-146 / dop853.rs:133-150)
            #[allow(clippy::too_many_arguments)]
            pub fn from_param(f: DeviceSystem, x: f64, x_end: f64, dx: f64, y: SVector<f64, D>, rtol: f64, atol: f64,
                              safety_factor: f64, beta: f64, fac_min: f64, fac_max: f64, h_max: f64, h: f64,
                              n_max: u32, n_stiff: u32, out_type: OutputType) -> Self {
                let mut cfg = ffi::ode_b200_config::default();
                unsafe {
                    ffi::ode_b200_config_from_param($method, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min,
                                                    fac_max, h_max, h, n_max, n_stiff, out_type as i32, &mut cfg)
                };
                Self { f, cfg, y, results: SolverResult::default(), stats: Stats::default() }
            }

            fn run(&mut self, ctx: &Context, mut com: Option<&mut ContinuousOutputModel<D>>)
                -> Result<Stats, IntegrationError> {
                let mut cap: usize = if self.cfg.out_type == ffi::ODE_B200_OUT_DENSE && self.cfg.dx > 0.0 {
                    ((self.cfg.x_end - self.cfg.x0) / self.cfg.dx).max(0.0) as usize + 3
                } else { 4096 };
                loop {
                    let (mut ox, mut oy) = (vec![0.0; cap], vec![0.0; cap * D]);
                    let (mut cb, mut cs, mut cc) = if com.is_some() {
                        (vec![0.0; cap], vec![0.0; cap], vec![0.0; cap * $m * D])
                    } else { (vec![], vec![], vec![]) };
                    let (mut yf, mut xf, mut hn) = (vec![0.0; D], [0.0f64], [0.0f64]);
                    let (mut ne, mut ac, mut rj, mut ns, mut st) = ([0u32], [0u32], [0u32], [0u32], [-1i32]);
                    let (mut oc, mut cn, mut cl) = ([0u32], [0u32], [0.0f64]);
                    let out = ffi::ode_b200_outputs {
                        y_final: yf.as_mut_ptr(), x_final: xf.as_mut_ptr(), h_next: hn.as_mut_ptr(),
                        num_eval: ne.as_mut_ptr(), accepted: ac.as_mut_ptr(), rejected: rj.as_mut_ptr(),
                        n_step: ns.as_mut_ptr(), status: st.as_mut_ptr(), out_cap: cap as i64,
                        out_x: ox.as_mut_ptr(), out_y: oy.as_mut_ptr(), out_count: oc.as_mut_ptr(),
                        com_cap: if com.is_some() { cap as i64 } else { 0 }, com_lower: cl.as_mut_ptr(),
                        com_break: cb.as_mut_ptr(), com_step: cs.as_mut_ptr(), com_coef: cc.as_mut_ptr(),
                        com_count: cn.as_mut_ptr(),
                    };
                    let rc = unsafe {
                        ffi::ode_b200_integrate_batch(ctx.0, self.f.sys_id, &self.cfg, 1, self.y.as_ptr(),
                                                      self.f.params.as_ptr(), 0, &out)
                    };
                    if rc != 0 { return Err(IntegrationError::Backend(last_error())); }
                    if oc[0] as usize > cap || cn[0] as usize > cap { cap = (oc[0].max(cn[0])) as usize; continue; }
                    let n = oc[0] as usize;
                    self.results = SolverResult(ox[..n].to_vec(),
                        (0..n).map(|s| SVector::<f64, D>::from_column_slice(&oy[s * D..(s + 1) * D])).collect());
                    self.stats = Stats { num_eval: ne[0], accepted_steps: ac[0], rejected_steps: rj[0] };
                    if let Some(model) = com.as_deref_mut() {
                        model.lower_bound = cl[0];
                        for q in 0..cn[0] as usize {
                            model.breakpoints.push(cb[q]);
                            model.intervals.push(Interval {
                                coefficients: (0..$m).map(|r| SVector::<f64, D>::from_column_slice(
                                    &cc[(q * $m + r) * D..(q * $m + r + 1) * D])).collect(),
                                step_size: cs[q],
                            });
                        }
                    }
                    return match st[0] {
                        ffi::ODE_B200_MAX_NUM_STEP_REACHED => Err(IntegrationError::MaxNumStepReached { x: xf[0], n_step: ns[0] }),
                        ffi::ODE_B200_STEP_SIZE_UNDERFLOW => Err(IntegrationError::StepSizeUnderflow { x: xf[0] }),
                        ffi::ODE_B200_STIFFNESS_DETECTED => Err(IntegrationError::StiffnessDetected { x: xf[0] }),
                        ffi::ODE_B200_UNSUPPORTED_DOMAIN => Err(IntegrationError::UnsupportedDomain),
                        _ => Ok(self.stats),
                    };
                }
            }

            pub fn x_out(&self) -> &Vec<f64> { self.results.get().0 }
            pub fn y_out(&self) -> &Vec<SVector<f64, D>> { self.results.get().1 }
            pub fn results(&self) -> &SolverResult<D> { &self.results }
        }

        impl<const D: usize> From<$name<D>> for SolverResult<D> {
            fn from(val: $name<D>) -> Self { val.results }
        }
    };
}

stepper!(Dopri5, ffi::ODE_B200_DOPRI5, 5);
stepper!(Dop853, ffi::ODE_B200_DOP853, 8);

impl<const D: usize> Dopri5<D> {
    /// src/dopri5.rs:255-260 (panics for OutputType::Continuous like the reference)
    pub fn integrate(&mut self, ctx: &Context) -> Result<Stats, IntegrationError> {
        if self.cfg.out_type == ffi::ODE_B200_OUT_CONTINUOUS {
            panic!("Please use `integrate_with_continuous_output_model` to compute a continuous output model.");
        }
        self.run(ctx, None)
    }
    /// src/dopri5.rs:246-252
    pub fn integrate_with_continuous_output_model(&mut self, ctx: &Context, model: &mut ContinuousOutputModel<D>)
        -> Result<Stats, IntegrationError> {
        self.cfg.out_type = ffi::ODE_B200_OUT_CONTINUOUS;
        self.run(ctx, Some(model))
    }
}

impl<const D: usize> Dop853<D> {
    /// src/dop853.rs:254
    pub fn integrate(&mut self, ctx: &Context) -> Result<Stats, IntegrationError> {
        self.run(ctx, None)
    }
}

/// src/rk4.rs -- `Rk4::new(f, x, y, x_end, step_size)` (note the argument order)
pub struct Rk4<const D: usize> {
    inner: Dop853<D>,
}
impl<const D: usize> Rk4<D> {
    pub fn new(f: DeviceSystem, x: f64, y: SVector<f64, D>, x_end: f64, step_size: f64) -> Self {
        let mut cfg = ffi::ode_b200_config::default();
        unsafe { ffi::ode_b200_config_rk4(x, x_end, step_size, &mut cfg) };
        Self { inner: Dop853 { f, cfg, y, results: SolverResult::default(), stats: Stats::default() } }
    }
    pub fn integrate(&mut self, ctx: &Context) -> Result<Stats, IntegrationError> { self.inner.run(ctx, None) }
    pub fn x_out(&self) -> &Vec<f64> { self.inner.x_out() }
    pub fn y_out(&self) -> &Vec<SVector<f64, D>> { self.inner.y_out() }
    pub fn results(&self) -> &SolverResult<D> { self.inner.results() }
}
