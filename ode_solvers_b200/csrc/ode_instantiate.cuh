/*
 * ode_instantiate.cuh -- instantiate the step-loop kernels of one system ahead of time.
 * One translation unit per built-in system (inst_*.cu) so `make -j` compiles them in parallel.
 * Dop853 ignores OutputType::Continuous (it behaves like Sparse: dop853.rs:398,540-542);
 * ODE_B200_OUT_CONTINUOUS8 is this library's Dop853-only extension.
 */
#pragma once

#include "ode_kernels.cuh"
#include "ode_registry.h"
#include "ode_systems.cuh"

#include "ode_f32.cuh"

namespace ode_b200 {
template <class Sys, int METHOD>
const void* f32_kernel_ptr() {
    if constexpr (Sys::HAS_F32)
        return (const void*)&ode_f32_step_loop_kernel<Sys, METHOD>;
    else
        return nullptr;
}
}  // namespace ode_b200
#define ODE_B200_F32_KSET(SYS)                                                                                \
    {ode_b200::f32_kernel_ptr<ode_b200::SYS, ODE_B200_RK4>(), ode_b200::f32_kernel_ptr<ode_b200::SYS, ODE_B200_DOPRI5>(), \
     ode_b200::f32_kernel_ptr<ode_b200::SYS, ODE_B200_DOP853>()}

#define ODE_B200_KPTR(...) ((const void*)&ode_b200::ode_step_loop_kernel<__VA_ARGS__>)

#define ODE_B200_KSET(SYS, ST)                                                                              \
    {{ODE_B200_KPTR(ode_b200::Rk4Stepper<ode_b200::SYS, ST>), ODE_B200_KPTR(ode_b200::Rk4Stepper<ode_b200::SYS, ST>), \
      ODE_B200_KPTR(ode_b200::Rk4Stepper<ode_b200::SYS, ST>), ODE_B200_KPTR(ode_b200::Rk4Stepper<ode_b200::SYS, ST>)}, \
     {ODE_B200_KPTR(ode_b200::Dopri5Stepper<ode_b200::SYS, ODE_B200_OUT_DENSE, ST>),                         \
      ODE_B200_KPTR(ode_b200::Dopri5Stepper<ode_b200::SYS, ODE_B200_OUT_SPARSE, ST>),                        \
      ODE_B200_KPTR(ode_b200::Dopri5Stepper<ode_b200::SYS, ODE_B200_OUT_CONTINUOUS, ST>), nullptr},          \
     {ODE_B200_KPTR(ode_b200::Dop853Stepper<ode_b200::SYS, ODE_B200_OUT_DENSE, ST>),                         \
      ODE_B200_KPTR(ode_b200::Dop853Stepper<ode_b200::SYS, ODE_B200_OUT_SPARSE, ST>),                        \
      ODE_B200_KPTR(ode_b200::Dop853Stepper<ode_b200::SYS, ODE_B200_OUT_SPARSE, ST>),                        \
      ODE_B200_KPTR(ode_b200::Dop853Stepper<ode_b200::SYS, ODE_B200_OUT_CONTINUOUS8, ST>)}}

#define ODE_B200_PKPTR(...) ((const void*)&ode_b200::ode_step_loop_kernel<__VA_ARGS__, true>)
#define ODE_B200_PIPE_KSET(SYS)                                                                              \
    {{ODE_B200_PKPTR(ode_b200::Rk4Stepper<ode_b200::SYS, false>), ODE_B200_PKPTR(ode_b200::Rk4Stepper<ode_b200::SYS, false>), \
      ODE_B200_PKPTR(ode_b200::Rk4Stepper<ode_b200::SYS, false>), ODE_B200_PKPTR(ode_b200::Rk4Stepper<ode_b200::SYS, false>)}, \
     {ODE_B200_PKPTR(ode_b200::Dopri5Stepper<ode_b200::SYS, ODE_B200_OUT_DENSE, false>),                      \
      ODE_B200_PKPTR(ode_b200::Dopri5Stepper<ode_b200::SYS, ODE_B200_OUT_SPARSE, false>),                     \
      ODE_B200_PKPTR(ode_b200::Dopri5Stepper<ode_b200::SYS, ODE_B200_OUT_CONTINUOUS, false>), nullptr},       \
     {ODE_B200_PKPTR(ode_b200::Dop853Stepper<ode_b200::SYS, ODE_B200_OUT_DENSE, false>),                      \
      ODE_B200_PKPTR(ode_b200::Dop853Stepper<ode_b200::SYS, ODE_B200_OUT_SPARSE, false>),                     \
      ODE_B200_PKPTR(ode_b200::Dop853Stepper<ode_b200::SYS, ODE_B200_OUT_SPARSE, false>),                     \
      ODE_B200_PKPTR(ode_b200::Dop853Stepper<ode_b200::SYS, ODE_B200_OUT_CONTINUOUS8, false>)}}

#define ODE_B200_NO_SPLIT \
    {{{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}}, {0, 0, 0, 0}, 0, {nullptr, nullptr, nullptr, nullptr}}
#define ODE_B200_DEFINE_SYSTEM(SYS, NAME)                                                \
    extern "C" const ode_b200::SystemEntry ode_b200_entry_##NAME = {                     \
        #NAME, ode_b200::SYS::DIM, ode_b200::SYS::NP, ode_b200::SYS::HAS_SOLOUT ? 1 : 0, \
        {ODE_B200_KSET(SYS, false), ODE_B200_KSET(SYS, true)}, ODE_B200_PIPE_KSET(SYS), ODE_B200_F32_KSET(SYS), \
        {ODE_B200_NO_SPLIT, ODE_B200_NO_SPLIT}, 0};

/* a system that also has a three-lanes-per-trajectory form (ode_split.cuh) for Dopri5 and Dop853 */
#define ODE_B200_SPLIT_KPTR(STEPPER, SPLIT, OUT, ST) \
    ((const void*)&ode_b200::ode_split_step_loop_kernel<ode_b200::STEPPER<ode_b200::SPLIT, OUT, ST>>)
#define ODE_B200_SPLIT_PKPTR(STEPPER, SPLIT, OUT) \
    ((const void*)&ode_b200::ode_split_step_loop_kernel<ode_b200::STEPPER<ode_b200::SPLIT, OUT, false>, true>)
#define ODE_B200_SPLIT_THREADS(STEPPER, SPLIT, OUT) ode_b200::STEPPER<ode_b200::SPLIT, OUT>::kThreads
/* O2 / O3: what the out_type slots Continuous / Continuous8 run (Dop853 treats Continuous like Sparse) */
#define ODE_B200_SPLIT_KSET(STEPPER, SPLIT, ST, O2)                                                                          \
    {ODE_B200_SPLIT_KPTR(STEPPER, SPLIT, ODE_B200_OUT_DENSE, ST), ODE_B200_SPLIT_KPTR(STEPPER, SPLIT, ODE_B200_OUT_SPARSE, ST), \
     ODE_B200_SPLIT_KPTR(STEPPER, SPLIT, O2, ST)
#define ODE_B200_SPLIT_ENTRY_DP5(SPLIT)                                                                                   \
    {{ODE_B200_SPLIT_KSET(Dopri5SplitStepper, SPLIT, false, ODE_B200_OUT_CONTINUOUS), nullptr},                           \
      ODE_B200_SPLIT_KSET(Dopri5SplitStepper, SPLIT, true, ODE_B200_OUT_CONTINUOUS), nullptr}},                           \
     {ODE_B200_SPLIT_THREADS(Dopri5SplitStepper, SPLIT, ODE_B200_OUT_DENSE),                                              \
      ODE_B200_SPLIT_THREADS(Dopri5SplitStepper, SPLIT, ODE_B200_OUT_SPARSE),                                             \
      ODE_B200_SPLIT_THREADS(Dopri5SplitStepper, SPLIT, ODE_B200_OUT_CONTINUOUS), 0},                                     \
     ode_b200::Dopri5SplitStepper<ode_b200::SPLIT, ODE_B200_OUT_SPARSE>::kSlotDoubles * 8,                                \
     {ODE_B200_SPLIT_PKPTR(Dopri5SplitStepper, SPLIT, ODE_B200_OUT_DENSE),                                                \
      ODE_B200_SPLIT_PKPTR(Dopri5SplitStepper, SPLIT, ODE_B200_OUT_SPARSE),                                               \
      ODE_B200_SPLIT_PKPTR(Dopri5SplitStepper, SPLIT, ODE_B200_OUT_CONTINUOUS), nullptr}}
#define ODE_B200_SPLIT_ENTRY_DP8(SPLIT)                                                                                   \
    {{ODE_B200_SPLIT_KSET(Dop853SplitStepper, SPLIT, false, ODE_B200_OUT_SPARSE),                                         \
      ODE_B200_SPLIT_KPTR(Dop853SplitStepper, SPLIT, ODE_B200_OUT_CONTINUOUS8, false)},                                   \
      ODE_B200_SPLIT_KSET(Dop853SplitStepper, SPLIT, true, ODE_B200_OUT_SPARSE),                                          \
      ODE_B200_SPLIT_KPTR(Dop853SplitStepper, SPLIT, ODE_B200_OUT_CONTINUOUS8, true)}},                                   \
     {ODE_B200_SPLIT_THREADS(Dop853SplitStepper, SPLIT, ODE_B200_OUT_DENSE),                                              \
      ODE_B200_SPLIT_THREADS(Dop853SplitStepper, SPLIT, ODE_B200_OUT_SPARSE),                                             \
      ODE_B200_SPLIT_THREADS(Dop853SplitStepper, SPLIT, ODE_B200_OUT_SPARSE),                                             \
      ODE_B200_SPLIT_THREADS(Dop853SplitStepper, SPLIT, ODE_B200_OUT_CONTINUOUS8)},                                       \
     ode_b200::Dop853SplitStepper<ode_b200::SPLIT, ODE_B200_OUT_SPARSE>::kSlotDoubles * 8,                                \
     {ODE_B200_SPLIT_PKPTR(Dop853SplitStepper, SPLIT, ODE_B200_OUT_DENSE),                                                \
      ODE_B200_SPLIT_PKPTR(Dop853SplitStepper, SPLIT, ODE_B200_OUT_SPARSE),                                               \
      ODE_B200_SPLIT_PKPTR(Dop853SplitStepper, SPLIT, ODE_B200_OUT_SPARSE),                                               \
      ODE_B200_SPLIT_PKPTR(Dop853SplitStepper, SPLIT, ODE_B200_OUT_CONTINUOUS8)}}
#define ODE_B200_DEFINE_SYSTEM_WITH_SPLIT(SYS, NAME, SPLIT)                                                   \
    extern "C" const ode_b200::SystemEntry ode_b200_entry_##NAME = {                                          \
        #NAME, ode_b200::SYS::DIM, ode_b200::SYS::NP, ode_b200::SYS::HAS_SOLOUT ? 1 : 0,                      \
        {ODE_B200_KSET(SYS, false), ODE_B200_KSET(SYS, true)}, ODE_B200_PIPE_KSET(SYS), ODE_B200_F32_KSET(SYS), \
        {ODE_B200_SPLIT_ENTRY_DP5(SPLIT), ODE_B200_SPLIT_ENTRY_DP8(SPLIT)}, 10};
