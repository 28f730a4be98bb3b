// Compile-check (and, on a GPU box, run) of the C++ host mirror ode_solvers_b200/cpp/ode_b200.hpp,
// written like the reference's src/dopri5.rs:545-556 test.
#include <cmath>
#include <cstdio>
#include "ode_b200.hpp"

using namespace ode_b200_host;

int main() {
    if (ode_b200_device_count() == 0) {
        std::printf("no gpu: compile/link check only\n");
        return 0;
    }
    Context ctx(0);
    auto s = Dopri5<1>::new_(System::builtin("test_exp_stop", {0.5}), 0.0, 1.0, 0.1, {1.0}, 1e-12, 1e-6);
    Stats st = s.integrate(ctx);
    const bool ok = std::fabs(s.x_out().back() - 0.5) < 1e-9 && s.y_out()[5][0] == 0.9130611474392001;
    std::printf("dopri5 test1: y_out[5] = %.16g accepted %u -> %s\n", s.y_out()[5][0], st.accepted_steps, ok ? "ok" : "FAIL");
    auto r = Rk4<1>::new_(System::builtin("test_exp"), 0.0, {1.0}, 1.0, 0.1);
    r.integrate(ctx);
    const bool ok2 = r.y_out().size() == 11 && std::fabs(r.y_out()[10][0] - 1.0715783953) < 1e-9;
    ContinuousOutputModel<1> com;
    auto c = Dopri5<1>::new_(System::builtin("test_cubic"), 0.0, 6.0, 0.1, {0.0}, 1e-10, 1e-10);
    c.integrate_with_continuous_output_model(ctx, com);
    auto v = com.evaluate(ctx, 3.0);
    const bool ok3 = v && std::fabs((*v)[0] + std::log(28.0)) < 1e-9 && !com.evaluate(ctx, 6.0001);
    std::printf("rk4 %s, continuous model %s\n", ok2 ? "ok" : "FAIL", ok3 ? "ok" : "FAIL");
    return ok && ok2 && ok3 ? 0 : 1;
}
