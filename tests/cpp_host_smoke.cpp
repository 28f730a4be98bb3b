// Compile-check (and, on a GPU box, run) of the C++ host mirror ode_solvers_b200/cpp/ode_b200.hpp,
// written like the reference's own tests (src/dopri5.rs:545-556, src/rk4.rs:150-170) and examples.
#include <cmath>
#include <cstdio>
#include <sstream>
#include "ode_b200.hpp"

using namespace ode_b200_host;

static int fails = 0;
static void report(const char* what, bool ok) {
    std::printf("%s: %s\n", what, ok ? "ok" : "FAIL");
    if (!ok) ++fails;
}

static int run();
int main() {
    try {
        return run();
    } catch (const std::exception& e) {
        std::printf("exception: %s -> FAIL\n", e.what());
        return 1;
    }
}

static int run() {
    // host-only parts of the surface
    report("abi version", abi_version() == ODE_B200_ABI_VERSION);
    report("tableau c5[1] = 1/5", tableau("c5", 1) == 1.0 / 5.0);
    {
        std::ostringstream os;
        os << Stats{11, 7, 1};
        report("Stats Display", os.str() == "Number of function evaluations: 11\nNumber of accepted steps: 7\nNumber of rejected steps: 1");
    }
    {
        ContinuousOutputModel<State<1, float>> cm;  // f32 models are evaluated on the host
        cm.lower_bound = 0.0f;
        cm.breakpoints = {1.0f};
        cm.intervals.push_back({{{1.0f}, {2.0f}}, 1.0f});
        auto v = cm.evaluate_host(0.5f);
        report("host evaluate (f32)", v && (*v)[0] == 2.0f && !cm.evaluate_host(-1.0f) && !cm.evaluate_host(1.5f));
    }
    System sys = System::builtin("lorenz", {10.0, 8.0 / 3.0, 28.0});
    report("system name / n_params", sys.name() == "lorenz" && sys.n_params() == 3 && sys.dim == 3);
    if (device_count() == 0) {
        std::printf("no gpu: compile/link check only (%d host checks failed)\n", fails);
        return fails ? 1 : 0;
    }
    Context ctx(0);
    // dopri5.rs:545-556 (test_integrate_test1_dopri5): stops at x = 0.5 through solout
    auto s = Dopri5<1>::new_(System::builtin("test_exp_stop", {0.5}), 0.0, 1.0, 0.1, {1.0}, 1e-12, 1e-6);
    Stats st = s.integrate(ctx);
    report("dopri5 test1", std::fabs(s.x_out().back() - 0.5) < 1e-9 && s.y_out()[5][0] == 0.9130611474392001 && st.accepted_steps > 0);
    // rk4.rs:150-170
    auto r = Rk4<1>::new_(System::builtin("test_exp"), 0.0, {1.0}, 1.0, 0.1);
    r.integrate(ctx);
    report("rk4", r.y_out().size() == 11 && std::fabs(r.y_out()[10][0] - 1.0715783953) < 1e-9);
    // continuous output model: Dopri5 (the crate's) and Dop853 (the Continuous8 extension)
    ContinuousOutputModel<State<1>> com;
    auto c = Dopri5<1>::new_(System::builtin("test_cubic"), 0.0, 6.0, 0.1, {0.0}, 1e-10, 1e-10);
    c.integrate_with_continuous_output_model(ctx, com);
    auto v = com.evaluate(ctx, 3.0);
    auto vh = com.evaluate_host(3.0);
    report("dopri5 continuous model", v && std::fabs((*v)[0] + std::log(28.0)) < 1e-9 && !com.evaluate(ctx, 6.0001) && vh &&
                                          std::fabs((*vh)[0] - (*v)[0]) < 1e-15);
    ContinuousOutputModel<State<3>> com8;
    auto c8 = Dop853<3>::new_(sys, 0.0, 1.0, 0.1, {1.0, 1.0, 1.0}, 1e-9, 1e-9);
    c8.integrate_with_continuous_output_model(ctx, com8);
    auto d8 = Dop853<3>::new_(sys, 0.0, 1.0, 0.1, {1.0, 1.0, 1.0}, 1e-9, 1e-9);  // Dense: samples at 0, 0.1, ..., 1.0
    d8.integrate(ctx);
    auto v8 = com8.evaluate(ctx, d8.x_out()[5]);
    report("dop853 continuous model", v8 && d8.y_out().size() == 11 && std::fabs((*v8)[0] - d8.y_out()[5][0]) < 1e-12 &&
                                          com8.intervals[0].coefficients.size() == 8 && !com8.evaluate(ctx, 1.5));
    // the *_dvector.rs examples: run-time state dimension, same bits as the SVector form
    auto kd = Dop853Dyn::new_(System::builtin("kepler", {398600.435436}), 0.0, 3600.0, 60.0,
                              DState<>{-5007.248417988539, -1444.918140151374, 3628.534606178356, 0.717716656891, -10.224093784269,
                                       0.748229399696},
                              1e-10, 1e-10);
    auto ks = Dop853<6>::new_(System::builtin("kepler", {398600.435436}), 0.0, 3600.0, 60.0,
                              {-5007.248417988539, -1444.918140151374, 3628.534606178356, 0.717716656891, -10.224093784269,
                               0.748229399696},
                              1e-10, 1e-10);
    Stats sd = kd.integrate(ctx), ss = ks.integrate(ctx);
    bool same = sd.num_eval == ss.num_eval && kd.y_out().size() == ks.y_out().size() && kd.y_out().size() == 61;
    for (size_t i = 0; same && i < kd.y_out().size(); ++i)
        for (int j = 0; j < 6; ++j) same = same && kd.y_out()[i][j] == ks.y_out()[i][j];
    report("DVector == SVector (Kepler, Dop853 dense)", same);
    // f32 (examples/bouncing_ball.rs:15,35-36): Rk4::<f32> stops at the ground through solout; Dopri5::<f32> runs
    auto b = Rk4<3, float>::new_(System::builtin("bouncing_ball", {9.81}), 0.0f, {10.0f, 0.0f, 0.0f}, 10.0f, 0.01f);
    b.integrate(ctx);
    report("rk4 f32 bouncing ball", !b.y_out().empty() && b.y_out().back()[0] < 0.0f && b.x_out().back() < 2.0f);
    auto b5 = Dopri5<3, float>::new_(System::builtin("bouncing_ball", {9.81}), 0.0f, 10.0f, 0.01f, {10.0f, 0.0f, 0.0f}, 1e-2f, 1e-6f);
    Stats s5 = b5.integrate(ctx);
    report("dopri5 f32 bouncing ball", s5.accepted_steps > 0 && !b5.y_out().empty());
    // batched call + knobs: 1000 Lorenz trajectories, with and without the work queue -> identical results
    const int n = 1000;
    std::vector<double> y0(3 * n);
    for (int i = 0; i < n; ++i) {
        y0[i] = -10.0 + 0.02 * i;
        y0[n + i] = 5.0;
        y0[2 * n + i] = 20.0;
    }
    ode_b200_config cfg;
    check(ode_b200_config_new(ODE_B200_DOPRI5, 0.0, 2.0, 0.1, 1e-6, 1e-6, &cfg), "config_new");
    cfg.out_type = ODE_B200_OUT_SPARSE;
    BatchResult r1 = integrate_batch(ctx, sys, cfg, n, y0.data());
    ctx.set_work_queue(false);
    ctx.set_pipeline(0);
    BatchResult r2 = integrate_batch(ctx, sys, cfg, n, y0.data());
    ctx.set_work_queue(true);
    bool eq = ctx.launch_count() >= 2 && ctx.last_kernel_ms() > 0.0 && ctx.stream() != nullptr;
    for (int i = 0; eq && i < n; ++i) eq = r1.status[i] == ODE_B200_OK && r1.y_final[i] == r2.y_final[i] && r1.accepted[i] == r2.accepted[i];
    report("integrate_batch (queue on / off)", eq);
    std::vector<float> y0f(y0.begin(), y0.end());
    ode_b200_config_f32 cf;
    check(ode_b200_config_new_f32(ODE_B200_DOP853, 0.0f, 2.0f, 0.1f, 1e-4f, 1e-4f, &cf), "config_new_f32");
    cf.out_type = ODE_B200_OUT_SPARSE;
    BatchResultF32 rf = integrate_batch_f32(ctx, sys, cf, n, y0f.data());
    BatchResultF32 rk = rk4_f32_integrate_batch(ctx, sys, 0.0f, 1.0f, 0.01f, n, y0f.data());
    report("integrate_batch_f32 / rk4_f32", rf.status[0] == ODE_B200_OK && rk.status[n - 1] == ODE_B200_OK && rk.accepted[0] == 100);
    // batched ContinuousOutputModel::evaluate of the model above at many points == its scalar evaluate()
    {
        const int64_t k = (int64_t)com.intervals.size();
        std::vector<double> stp(k), coef((size_t)k * 5);
        for (int64_t q = 0; q < k; ++q) {
            stp[q] = com.intervals[q].step_size;
            for (int rr = 0; rr < 5; ++rr) coef[(size_t)q * 5 + rr] = com.intervals[q].coefficients[rr][0];
        }
        uint32_t cnt = (uint32_t)k;
        auto res = com_evaluate_batch(ctx, 1, 1, 5, k, &com.lower_bound, com.breakpoints.data(), stp.data(), coef.data(), &cnt,
                                      {0.5, 3.0, 5.5, 7.0});
        report("com_evaluate_batch", res.second[0] && res.second[1] && res.second[2] && !res.second[3] && res.first[1] == (*v)[0]);
    }
    std::printf("%d checks failed\n", fails);
    return fails ? 1 : 0;
}
