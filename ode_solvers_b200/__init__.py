"""ode_solvers_b200 -- B200-native batched explicit Runge-Kutta step loop (Rk4, Dopri5, Dop853)
behind the API surface of the Rust crate `ode_solvers`. All numerics run in
libode_b200.so (hand-written sm_100a CUDA); there is no CPU fallback."""
from . import _abi, workloads  # noqa: F401
from ._abi import (DOP853, DOPRI5, RK4, OUT_CONTINUOUS, OUT_CONTINUOUS8, OUT_DENSE, OUT_SPARSE,  # noqa: F401
                   OK, MAX_NUM_STEP_REACHED, STEP_SIZE_UNDERFLOW, STIFFNESS_DETECTED,
                   OUTPUT_CAPACITY_EXCEEDED, UNSUPPORTED_DOMAIN)
from .api import (Context, ContinuousOutputModel, Dop853, Dopri5, DynSystem, IntegrationError,  # noqa: F401
                  MaxNumStepReached, Ode200Error, OutputType, Rk4, SolverResult, Stats,
                  StepSizeUnderflow, StiffnessDetected, System, UnsupportedDomain, com_evaluate_batch,
                  config_from_param, config_new, config_rk4, default_context, device_count,
                  integrate_batch, integrate_batch_f32, config_new_f32, config_from_param_f32, config_rk4_f32,
                  measure_fp64_peak, rk4_f32_integrate_batch, save)
