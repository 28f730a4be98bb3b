// ode_b200.hpp -- header-only C++17 mirror of the reference crate's surface over the C ABI.
//
// The reference is compiled code (Rust) whose toolchain is absent from this image, so the
// compiled-language host layer above the ABI is provided in C++ with the crate's names and
// argument order: System (trait System, src/dop_shared.rs:31-42), Dopri5::new_/from_param
// (src/dopri5.rs:76,129), Dop853 (src/dop853.rs:76,133), Rk4::new_(f, x, y, x_end, step)
// (src/rk4.rs:42), integrate(), integrate_with_continuous_output_model(), x_out()/y_out(),
// Stats, IntegrationError, ContinuousOutputModel::evaluate/bounds.
// All numerics run in libode_b200.so; nothing here computes a step.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "ode_b200.h"

namespace ode_b200_host {

enum class OutputType : int { Dense = ODE_B200_OUT_DENSE, Sparse = ODE_B200_OUT_SPARSE, Continuous = ODE_B200_OUT_CONTINUOUS };

struct Stats {  // src/dop_shared.rs:131-136
    uint32_t num_eval = 0, accepted_steps = 0, rejected_steps = 0;
};

struct IntegrationError : std::runtime_error {  // src/dop_shared.rs:120-128
    int status;
    double x;
    uint32_t n_step;
    IntegrationError(int st, double x_, uint32_t n)
        : std::runtime_error(st == ODE_B200_MAX_NUM_STEP_REACHED
                                 ? "Stopped at x = " + std::to_string(x_) + ". Need more than " + std::to_string(n) + " steps."
                             : st == ODE_B200_STEP_SIZE_UNDERFLOW ? "Stopped at x = " + std::to_string(x_) + ". Step size underflow."
                             : st == ODE_B200_STIFFNESS_DETECTED  ? "The problem seems to become stiff at x = " + std::to_string(x_) + "."
                                                                  : "input outside the domain the reference handles"),
          status(st), x(x_), n_step(n) {}
};

struct BackendError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void check(int rc, const char* what) {
    if (rc != 0) throw BackendError(std::string(what) + ": " + ode_b200_last_error());
}

class Context {
   public:
    explicit Context(int device = 0) : p_(ode_b200_create(device)) {
        if (!p_) throw BackendError(std::string("ode_b200_create: ") + ode_b200_last_error());
    }
    ~Context() { ode_b200_destroy(p_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    ode_b200_ctx* get() const { return p_; }

   private:
    ode_b200_ctx* p_;
};

// trait System: a device functor (built in or NVRTC source) + the struct fields as parameters
struct System {
    int sys_id = -1;
    int dim = 0;
    std::vector<double> params;

    static System builtin(const std::string& name, std::vector<double> params = {}) {
        System s;
        s.sys_id = ode_b200_system_builtin(name.c_str());
        if (s.sys_id < 0) throw BackendError(ode_b200_last_error());
        s.dim = ode_b200_system_dim(s.sys_id);
        s.params = std::move(params);
        return s;
    }
    static System from_source(const std::string& name, int dim, const std::string& rhs_body,
                              const std::string& solout_body = "", std::vector<double> params = {}) {
        System s;
        s.sys_id = ode_b200_system_from_source(name.c_str(), dim, (int)params.size(), rhs_body.c_str(),
                                               solout_body.empty() ? nullptr : solout_body.c_str());
        if (s.sys_id < 0) throw BackendError(ode_b200_last_error());
        s.dim = dim;
        s.params = std::move(params);
        return s;
    }
};

template <int D>
using State = std::array<double, D>;

// src/continuous_output_model.rs:8-104
template <int D>
struct ContinuousOutputModel {
    struct Interval {
        std::vector<State<D>> coefficients;
        double step_size;
    };
    double lower_bound = 0.0;
    std::vector<double> breakpoints;
    std::vector<Interval> intervals;

    std::pair<double, double> bounds() const { return {lower_bound, breakpoints.back()}; }

    // evaluate() on the device through the batched kernel (N = 1)
    std::optional<State<D>> evaluate(const Context& ctx, double x) const {
        if (intervals.empty()) return std::nullopt;
        const int m = (int)intervals[0].coefficients.size();
        const int64_t k = (int64_t)intervals.size();
        std::vector<double> brk(breakpoints), stp(k), coef((size_t)k * m * D);
        for (int64_t q = 0; q < k; ++q) {
            stp[q] = intervals[q].step_size;
            for (int r = 0; r < m; ++r)
                for (int c = 0; c < D; ++c) coef[((size_t)q * m + r) * D + c] = intervals[q].coefficients[r][c];
        }
        uint32_t count = (uint32_t)k;
        State<D> y{};
        uint8_t found = 0;
        check(ode_b200_com_evaluate(ctx.get(), 1, D, m, k, &lower_bound, brk.data(), stp.data(), coef.data(), &count, 1,
                                    &x, y.data(), &found),
              "ode_b200_com_evaluate");
        if (!found) return std::nullopt;
        return y;
    }
};

template <int D>
struct SolverResult {  // src/dop_shared.rs:46,79-109
    std::vector<double> x;
    std::vector<State<D>> y;
    void append(SolverResult&& o) {
        x.insert(x.end(), o.x.begin(), o.x.end());
        y.insert(y.end(), o.y.begin(), o.y.end());
    }
};

template <int D, int METHOD>
class Stepper {
   public:
    Stats integrate(const Context& ctx) {
        if (METHOD == ODE_B200_DOPRI5 && cfg_.out_type == ODE_B200_OUT_CONTINUOUS)  // dopri5.rs:256-258 panics
            throw std::logic_error("Please use `integrate_with_continuous_output_model` to compute a continuous output model.");
        return run(ctx, nullptr);
    }
    const std::vector<double>& x_out() const { return results_.x; }
    const std::vector<State<D>>& y_out() const { return results_.y; }
    const SolverResult<D>& results() const { return results_; }
    SolverResult<D> into() && { return std::move(results_); }
    ode_b200_config& config() { return cfg_; }

   protected:
    Stepper(System f, const ode_b200_config& cfg, const State<D>& y) : f_(std::move(f)), cfg_(cfg), y_(y) {
        if (f_.dim != D) throw std::invalid_argument("state dimension does not match the system");
    }

    Stats run(const Context& ctx, ContinuousOutputModel<D>* com) {
        constexpr int M = METHOD == ODE_B200_DOPRI5 ? 5 : 8;
        int64_t cap = 4096;
        if (METHOD == ODE_B200_RK4 && cfg_.step_size != 0.0)
            cap = (int64_t)std::max(0.0, std::ceil((cfg_.x_end - cfg_.x0) / cfg_.step_size)) + 2;
        else if (cfg_.out_type == ODE_B200_OUT_DENSE && cfg_.dx > 0.0)
            cap = (int64_t)std::max(0.0, (cfg_.x_end - cfg_.x0) / cfg_.dx) + 3;
        for (;;) {
            std::vector<double> ox(cap), oy((size_t)cap * D), cb, cs, cc;
            if (com) {
                cb.resize(cap);
                cs.resize(cap);
                cc.resize((size_t)cap * M * D);
            }
            State<D> yf{};
            double xf = 0, hn = 0, cl = 0;
            uint32_t ne = 0, ac = 0, rj = 0, ns = 0, oc = 0, cn = 0;
            int32_t st = -1;
            ode_b200_outputs o{};
            o.y_final = yf.data(); o.x_final = &xf; o.h_next = &hn; o.num_eval = &ne; o.accepted = &ac;
            o.rejected = &rj; o.n_step = &ns; o.status = &st; o.out_cap = cap; o.out_x = ox.data();
            o.out_y = oy.data(); o.out_count = &oc; o.com_cap = com ? cap : 0; o.com_lower = &cl;
            o.com_break = cb.data(); o.com_step = cs.data(); o.com_coef = cc.data(); o.com_count = &cn;
            check(ode_b200_integrate_batch(ctx.get(), f_.sys_id, &cfg_, 1, y_.data(), f_.params.data(), 0, &o),
                  "ode_b200_integrate_batch");
            if ((int64_t)oc > cap || (int64_t)cn > cap) {
                cap = std::max<int64_t>(oc, cn);
                continue;
            }
            results_.x.assign(ox.begin(), ox.begin() + oc);
            results_.y.resize(oc);
            for (uint32_t s = 0; s < oc; ++s) std::copy_n(&oy[(size_t)s * D], D, results_.y[s].begin());
            stats_ = Stats{ne, ac, rj};
            if (com) {
                com->lower_bound = cl;
                for (uint32_t q = 0; q < cn; ++q) {
                    typename ContinuousOutputModel<D>::Interval iv;
                    iv.step_size = cs[q];
                    iv.coefficients.resize(M);
                    for (int r = 0; r < M; ++r) std::copy_n(&cc[((size_t)q * M + r) * D], D, iv.coefficients[r].begin());
                    com->breakpoints.push_back(cb[q]);
                    com->intervals.push_back(std::move(iv));
                }
            }
            if (st != ODE_B200_OK && st != ODE_B200_OUTPUT_CAPACITY_EXCEEDED) throw IntegrationError(st, xf, ns);
            return stats_;
        }
    }

    System f_;
    ode_b200_config cfg_;
    State<D> y_;
    SolverResult<D> results_;
    Stats stats_;
};

template <int D>
class Dopri5 : public Stepper<D, ODE_B200_DOPRI5> {
    using Base = Stepper<D, ODE_B200_DOPRI5>;
    Dopri5(System f, const ode_b200_config& c, const State<D>& y) : Base(std::move(f), c, y) {}

   public:
    static Dopri5 new_(System f, double x, double x_end, double dx, const State<D>& y, double rtol, double atol) {
        ode_b200_config c;
        check(ode_b200_config_new(ODE_B200_DOPRI5, x, x_end, dx, rtol, atol, &c), "config_new");
        return Dopri5(std::move(f), c, y);
    }
    static Dopri5 from_param(System f, double x, double x_end, double dx, const State<D>& y, double rtol, double atol,
                             double safety_factor, double beta, double fac_min, double fac_max, double h_max, double h,
                             uint32_t n_max, uint32_t n_stiff, OutputType out_type) {
        ode_b200_config c;
        check(ode_b200_config_from_param(ODE_B200_DOPRI5, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min,
                                         fac_max, h_max, h, n_max, n_stiff, (int)out_type, &c),
              "config_from_param");
        return Dopri5(std::move(f), c, y);
    }
    Stats integrate_with_continuous_output_model(const Context& ctx, ContinuousOutputModel<D>& model) {
        this->cfg_.out_type = ODE_B200_OUT_CONTINUOUS;  // dopri5.rs:250
        return this->run(ctx, &model);
    }
};

template <int D>
class Dop853 : public Stepper<D, ODE_B200_DOP853> {
    using Base = Stepper<D, ODE_B200_DOP853>;
    Dop853(System f, const ode_b200_config& c, const State<D>& y) : Base(std::move(f), c, y) {}

   public:
    static Dop853 new_(System f, double x, double x_end, double dx, const State<D>& y, double rtol, double atol) {
        ode_b200_config c;
        check(ode_b200_config_new(ODE_B200_DOP853, x, x_end, dx, rtol, atol, &c), "config_new");
        return Dop853(std::move(f), c, y);
    }
    static Dop853 from_param(System f, double x, double x_end, double dx, const State<D>& y, double rtol, double atol,
                             double safety_factor, double beta, double fac_min, double fac_max, double h_max, double h,
                             uint32_t n_max, uint32_t n_stiff, OutputType out_type) {
        ode_b200_config c;
        check(ode_b200_config_from_param(ODE_B200_DOP853, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min,
                                         fac_max, h_max, h, n_max, n_stiff, (int)out_type, &c),
              "config_from_param");
        return Dop853(std::move(f), c, y);
    }
};

template <int D>
class Rk4 : public Stepper<D, ODE_B200_RK4> {
    using Base = Stepper<D, ODE_B200_RK4>;
    Rk4(System f, const ode_b200_config& c, const State<D>& y) : Base(std::move(f), c, y) {}

   public:
    // Rk4::new(f, x, y, x_end, step_size) -- src/rk4.rs:42
    static Rk4 new_(System f, double x, const State<D>& y, double x_end, double step_size) {
        ode_b200_config c;
        check(ode_b200_config_rk4(x, x_end, step_size, &c), "config_rk4");
        return Rk4(std::move(f), c, y);
    }
};

}  // namespace ode_b200_host
