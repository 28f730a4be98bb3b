#include "ode_instantiate.cuh"
ODE_B200_DEFINE_SYSTEM(SysThreeBody18, threebody18)
