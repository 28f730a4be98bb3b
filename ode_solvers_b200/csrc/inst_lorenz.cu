#include "ode_instantiate.cuh"
ODE_B200_DEFINE_SYSTEM(SysLorenz, lorenz)
