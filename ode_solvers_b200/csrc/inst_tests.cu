/* the reference's own test ODEs (src/rk4.rs:178-221, src/dopri5.rs:531-543, tests/dopri5.rs:13-31) */
#include "ode_instantiate.cuh"
ODE_B200_DEFINE_SYSTEM(SysTestRk4_1, test_rk4_1)
ODE_B200_DEFINE_SYSTEM(SysTestRk4_2, test_rk4_2)
ODE_B200_DEFINE_SYSTEM(SysTestExp, test_exp)
ODE_B200_DEFINE_SYSTEM(SysTestExpStop, test_exp_stop)
ODE_B200_DEFINE_SYSTEM(SysTestCubic, test_cubic)
ODE_B200_DEFINE_SYSTEM(SysTestCubicStop, test_cubic_stop)
