//! Raw bindings of include/ode_b200.h (field order and types must match the C header).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

pub const ODE_B200_RK4: c_int = 0;
pub const ODE_B200_DOPRI5: c_int = 1;
pub const ODE_B200_DOP853: c_int = 2;

pub const ODE_B200_OUT_DENSE: c_int = 0;
pub const ODE_B200_OUT_SPARSE: c_int = 1;
pub const ODE_B200_OUT_CONTINUOUS: c_int = 2;

pub const ODE_B200_OK: i32 = 0;
pub const ODE_B200_MAX_NUM_STEP_REACHED: i32 = 1;
pub const ODE_B200_STEP_SIZE_UNDERFLOW: i32 = 2;
pub const ODE_B200_STIFFNESS_DETECTED: i32 = 3;
pub const ODE_B200_OUTPUT_CAPACITY_EXCEEDED: i32 = 4;
pub const ODE_B200_UNSUPPORTED_DOMAIN: i32 = 5;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ode_b200_config {
    pub method: i32,
    pub out_type: i32,
    pub x0: c_double,
    pub x_end: c_double,
    pub dx: c_double,
    pub rtol: c_double,
    pub atol: c_double,
    pub safety_factor: c_double,
    pub alpha: c_double,
    pub beta: c_double,
    pub fac_min: c_double,
    pub fac_max: c_double,
    pub h_max: c_double,
    pub h0: c_double,
    pub uround: c_double,
    pub step_size: c_double,
    pub n_max: u32,
    pub n_stiff: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct ode_b200_outputs {
    pub y_final: *mut c_double,
    pub x_final: *mut c_double,
    pub h_next: *mut c_double,
    pub num_eval: *mut u32,
    pub accepted: *mut u32,
    pub rejected: *mut u32,
    pub n_step: *mut u32,
    pub status: *mut i32,
    pub out_cap: i64,
    pub out_x: *mut c_double,
    pub out_y: *mut c_double,
    pub out_count: *mut u32,
    pub com_cap: i64,
    pub com_lower: *mut c_double,
    pub com_break: *mut c_double,
    pub com_step: *mut c_double,
    pub com_coef: *mut c_double,
    pub com_count: *mut u32,
}

#[repr(C)]
pub struct ode_b200_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn ode_b200_abi_version() -> c_int;
    pub fn ode_b200_device_count() -> c_int;
    pub fn ode_b200_last_error() -> *const c_char;
    pub fn ode_b200_system_builtin(name: *const c_char) -> c_int;
    pub fn ode_b200_system_dim(sys_id: c_int) -> c_int;
    pub fn ode_b200_system_n_params(sys_id: c_int) -> c_int;
    pub fn ode_b200_system_from_source(
        name: *const c_char, dim: c_int, n_params: c_int, rhs_body: *const c_char, solout_body: *const c_char,
    ) -> c_int;
    pub fn ode_b200_system_compile(sys_id: c_int, method: c_int, out_type: c_int) -> c_int;
    pub fn ode_b200_config_new(
        method: c_int, x: c_double, x_end: c_double, dx: c_double, rtol: c_double, atol: c_double,
        out: *mut ode_b200_config,
    ) -> c_int;
    pub fn ode_b200_config_from_param(
        method: c_int, x: c_double, x_end: c_double, dx: c_double, rtol: c_double, atol: c_double,
        safety_factor: c_double, beta: c_double, fac_min: c_double, fac_max: c_double, h_max: c_double,
        h: c_double, n_max: u32, n_stiff: u32, out_type: c_int, out: *mut ode_b200_config,
    ) -> c_int;
    pub fn ode_b200_config_rk4(x: c_double, x_end: c_double, step_size: c_double, out: *mut ode_b200_config) -> c_int;
    pub fn ode_b200_create(device: c_int) -> *mut ode_b200_ctx;
    pub fn ode_b200_destroy(ctx: *mut ode_b200_ctx);
    pub fn ode_b200_integrate_batch(
        ctx: *mut ode_b200_ctx, sys_id: c_int, cfg: *const ode_b200_config, n_traj: i64, y0: *const c_double,
        params: *const c_double, params_per_traj: c_int, out: *const ode_b200_outputs,
    ) -> c_int;
    pub fn ode_b200_integrate_batch_device(
        ctx: *mut ode_b200_ctx, sys_id: c_int, cfg: *const ode_b200_config, n_traj: i64, y0: *const c_double,
        params: *const c_double, params_per_traj: c_int, out: *const ode_b200_outputs, cuda_stream: *mut c_void,
    ) -> c_int;
    pub fn ode_b200_integrate_batch_multi(
        sys_id: c_int, cfg: *const ode_b200_config, n_traj: i64, y0: *const c_double, params: *const c_double,
        params_per_traj: c_int, n_gpus: c_int, out: *const ode_b200_outputs,
    ) -> c_int;
    /// 1 (default): persistent lanes refill from a work queue; 0: one thread per trajectory.
    pub fn ode_b200_set_work_queue(ctx: *mut ode_b200_ctx, enabled: c_int) -> c_int;
    /// 1 (default): per-trajectory results are stored straight into pinned host arrays; 0: device
    /// buffers + one DMA copy per array (hosts whose GPUs share PCIe with 7 others).
    pub fn ode_b200_set_zero_copy(ctx: *mut ode_b200_ctx, enabled: c_int) -> c_int;
    pub fn ode_b200_com_evaluate(
        ctx: *mut ode_b200_ctx, n_traj: i64, dim: c_int, m: c_int, com_cap: i64, com_lower: *const c_double,
        com_break: *const c_double, com_step: *const c_double, com_coef: *const c_double, com_count: *const u32,
        n_query: i64, xq: *const c_double, y: *mut c_double, found: *mut u8,
    ) -> c_int;
}
