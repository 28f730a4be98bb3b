#include "ode_instantiate.cuh"
#include "ode_split.cuh"
ODE_B200_DEFINE_SYSTEM_WITH_SPLIT(SysThreeBody18, threebody18, SplitThreeBody18)
