#include "ode_instantiate.cuh"
ODE_B200_DEFINE_SYSTEM(SysCr3bp, cr3bp)
