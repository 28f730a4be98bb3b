// ode_b200.hpp -- header-only C++17 mirror of the reference crate's surface over the C ABI.
//
// The reference is compiled code (Rust) whose toolchain is absent from this image, so the
// compiled-language host layer above the ABI is provided in C++ with the crate's names and
// argument order: System (trait System, src/dop_shared.rs:31-42), Dopri5::new_/from_param
// (src/dopri5.rs:76,129), Dop853 (src/dop853.rs:76,133), Rk4::new_(f, x, y, x_end, step)
// (src/rk4.rs:42), integrate(), integrate_with_continuous_output_model(), x_out()/y_out(),
// Stats (with its Display text), IntegrationError, SolverResult, ContinuousOutputModel::evaluate/bounds.
// Like the crate the steppers are generic over the scalar (f64 / f32: `impl FloatNumber for f32`,
// src/dop_shared.rs:74) and over the state vector: State<D, T> (= SVector<T, D>) or DState<T>
// (= DVector<T>, the *_dvector.rs examples). New with the batched ABI: integrate_batch (many
// trajectories, one configuration; one or several GPUs), the device-pointer entries and the batched
// ContinuousOutputModel evaluation. rust/src/lib.rs is the same surface in the reference's own language.
// All numerics run in libode_b200.so; nothing here computes a step.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <optional>
#include <ostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "ode_b200.h"

namespace ode_b200_host {

enum class OutputType : int { Dense = ODE_B200_OUT_DENSE, Sparse = ODE_B200_OUT_SPARSE, Continuous = ODE_B200_OUT_CONTINUOUS };

struct Stats {  // src/dop_shared.rs:131-153
    uint32_t num_eval = 0, accepted_steps = 0, rejected_steps = 0;
    friend std::ostream& operator<<(std::ostream& os, const Stats& s) {  // impl fmt::Display for Stats
        return os << "Number of function evaluations: " << s.num_eval << "\nNumber of accepted steps: " << s.accepted_steps
                  << "\nNumber of rejected steps: " << s.rejected_steps;
    }
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

inline int abi_version() { return ode_b200_abi_version(); }
inline int device_count() { return ode_b200_device_count(); }
// (FP64 thread-instructions/s of a DFMA chain microbenchmark, SM clock estimate in MHz) of one GPU: the roofline
// denominator bench.py reports against
inline std::pair<double, double> measure_fp64_peak(int device = 0) {
    double inst_per_s = 0, mhz = 0;
    check(ode_b200_measure_fp64_peak(device, &inst_per_s, &mhz), "ode_b200_measure_fp64_peak");
    return {inst_per_s, mhz};
}
// butcher_tableau.rs accessors as the kernels hold them ("a5","c5","d5","e5","a8","b8","bhh8","c8","d8","e8")
inline double tableau(const std::string& which, int i, int j = 0) { return ode_b200_tableau(which.c_str(), i, j); }

class Context {
   public:
    explicit Context(int device = 0) : p_(ode_b200_create(device)) {
        if (!p_) throw BackendError(std::string("ode_b200_create: ") + ode_b200_last_error());
    }
    ~Context() { ode_b200_destroy(p_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    ode_b200_ctx* get() const { return p_; }
    // transfer / scheduling knobs of the batched call (defaults: pipeline 2^16, zero-copy, work queue, lane split)
    void set_pipeline(int chunk_log2) const { check(ode_b200_set_pipeline(p_, chunk_log2), "ode_b200_set_pipeline"); }
    void set_zero_copy(bool on) const { check(ode_b200_set_zero_copy(p_, on ? 1 : 0), "ode_b200_set_zero_copy"); }
    void set_work_queue(bool on) const { check(ode_b200_set_work_queue(p_, on ? 1 : 0), "ode_b200_set_work_queue"); }
    void set_lane_split(bool on) const { check(ode_b200_set_lane_split(p_, on ? 1 : 0), "ode_b200_set_lane_split"); }
    double last_kernel_ms() const { return ode_b200_last_kernel_ms(p_); }
    uint64_t launch_count() const { return (uint64_t)ode_b200_launch_count(p_); }
    void* stream() const { return ode_b200_stream(p_); }  // the context's cudaStream_t

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
    std::string name() const { return ode_b200_system_name(sys_id); }
    int n_params() const { return ode_b200_system_n_params(sys_id); }
    // compile a source system's kernel for (method, out_type) ahead of the first call (built-ins: no-op)
    void compile(int method, OutputType out_type) const {
        check(ode_b200_system_compile(sys_id, method, (int)out_type), "ode_b200_system_compile");
    }
};

// ---- state vectors: State<D, T> ~ SVector<T, D>, DState<T> ~ DVector<T>
template <int D, class T = double>
using State = std::array<T, D>;
template <class T = double>
using DState = std::vector<T>;

namespace detail {
template <class V>
struct VecTraits;
template <class T, size_t D>
struct VecTraits<std::array<T, D>> {
    using Scalar = T;
    static size_t dim(const std::array<T, D>&) { return D; }
    static std::array<T, D> from_flat(const T* p, size_t) {
        std::array<T, D> v;
        std::copy_n(p, D, v.begin());
        return v;
    }
};
template <class T>
struct VecTraits<std::vector<T>> {
    using Scalar = T;
    static size_t dim(const std::vector<T>& v) { return v.size(); }
    static std::vector<T> from_flat(const T* p, size_t n) { return std::vector<T>(p, p + n); }
};

// the two instantiations of the ABI (include/ode_b200.h: ode_b200_config / ode_b200_config_f32, ...)
template <class T>
struct Abi;
template <>
struct Abi<double> {
    using Config = ode_b200_config;
    using Outputs = ode_b200_outputs;
    static int config_new(int m, double x, double xe, double dx, double rtol, double atol, Config* c) {
        return ode_b200_config_new(m, x, xe, dx, rtol, atol, c);
    }
    static int config_from_param(int m, double x, double xe, double dx, double rtol, double atol, double sf, double beta,
                                 double fmin, double fmax, double hmax, double h, uint32_t n_max, uint32_t n_stiff, int ot,
                                 Config* c) {
        return ode_b200_config_from_param(m, x, xe, dx, rtol, atol, sf, beta, fmin, fmax, hmax, h, n_max, n_stiff, ot, c);
    }
    static int config_rk4(double x, double xe, double step, Config* c) { return ode_b200_config_rk4(x, xe, step, c); }
    static int integrate(ode_b200_ctx* ctx, int sys, const Config* c, int64_t n, const double* y0, const double* p, int ppt,
                         const Outputs* o) {
        return ode_b200_integrate_batch(ctx, sys, c, n, y0, p, ppt, o);
    }
};
template <>
struct Abi<float> {
    using Config = ode_b200_config_f32;
    using Outputs = ode_b200_outputs_f32;
    static int config_new(int m, float x, float xe, float dx, float rtol, float atol, Config* c) {
        return ode_b200_config_new_f32(m, x, xe, dx, rtol, atol, c);
    }
    static int config_from_param(int m, float x, float xe, float dx, float rtol, float atol, float sf, float beta, float fmin,
                                 float fmax, float hmax, float h, uint32_t n_max, uint32_t n_stiff, int ot, Config* c) {
        return ode_b200_config_from_param_f32(m, x, xe, dx, rtol, atol, sf, beta, fmin, fmax, hmax, h, n_max, n_stiff, ot, c);
    }
    static int config_rk4(float x, float xe, float step, Config* c) { return ode_b200_config_rk4_f32(x, xe, step, c); }
    static int integrate(ode_b200_ctx* ctx, int sys, const Config* c, int64_t n, const float* y0, const float* p, int ppt,
                         const Outputs* o) {
        return ode_b200_integrate_batch_f32(ctx, sys, c, n, y0, p, ppt, o);
    }
};
}  // namespace detail

// src/continuous_output_model.rs:8-104 (field names as there)
template <class V>
struct ContinuousOutputModel {
    using T = typename detail::VecTraits<V>::Scalar;
    struct Interval {
        std::vector<V> coefficients;
        T step_size;
    };
    T lower_bound = 0;
    std::vector<T> breakpoints;
    std::vector<Interval> intervals;

    std::pair<T, T> bounds() const { return {lower_bound, breakpoints.back()}; }

    // evaluate() (continuous_output_model.rs:31-56,86-104) through the batched device kernel with N = 1
    // (f64; ode_b200_com_evaluate). For many models / many query points use com_evaluate_batch below.
    std::optional<V> evaluate(const Context& ctx, T x) const {
        if constexpr (!std::is_same<T, double>::value) {
            (void)ctx;
            return evaluate_host(x);
        } else {
            if (intervals.empty()) return std::nullopt;
            const int m = (int)intervals[0].coefficients.size();
            const size_t d = detail::VecTraits<V>::dim(intervals[0].coefficients[0]);
            const int64_t k = (int64_t)intervals.size();
            std::vector<double> brk(breakpoints), stp(k), coef((size_t)k * m * d), y(d);
            for (int64_t q = 0; q < k; ++q) {
                stp[q] = intervals[q].step_size;
                for (int r = 0; r < m; ++r)
                    for (size_t c = 0; c < d; ++c) coef[((size_t)q * m + r) * d + c] = intervals[q].coefficients[r][c];
            }
            uint32_t count = (uint32_t)k;
            uint8_t found = 0;
            check(ode_b200_com_evaluate(ctx.get(), 1, (int)d, m, k, &lower_bound, brk.data(), stp.data(), coef.data(), &count,
                                        1, &x, y.data(), &found),
                  "ode_b200_com_evaluate");
            if (!found) return std::nullopt;
            return detail::VecTraits<V>::from_flat(y.data(), d);
        }
    }
    // the same on the host in T arithmetic (one rounding per operation when the translation unit is compiled
    // without floating-point contraction, as the crate is)
    std::optional<V> evaluate_host(T x) const {
        if (x < lower_bound) return std::nullopt;
        const size_t idx = (size_t)(std::lower_bound(breakpoints.begin(), breakpoints.end(), x) - breakpoints.begin());
        if (idx >= intervals.size()) return std::nullopt;
        const T lower = idx == 0 ? lower_bound : breakpoints[idx - 1];
        const Interval& iv = intervals[idx];
        const T theta = (x - lower) / iv.step_size, theta1 = T(1) - theta;
        const std::vector<V>& c = iv.coefficients;
        V r = c.back();
        for (size_t i = c.size() - 1; i-- > 0;)
            for (size_t j = 0; j < detail::VecTraits<V>::dim(r); ++j) r[j] = c[i][j] + r[j] * (i % 2 == 0 ? theta : theta1);
        return r;
    }
};

template <class V>
struct SolverResult {  // src/dop_shared.rs:46,79-109
    using T = typename detail::VecTraits<V>::Scalar;
    std::vector<T> x;
    std::vector<V> y;
    void push(T xi, V yi) {
        x.push_back(xi);
        y.push_back(std::move(yi));
    }
    std::pair<const std::vector<T>&, const std::vector<V>&> get() const { return {x, y}; }
    void append(SolverResult&& o) {
        x.insert(x.end(), o.x.begin(), o.x.end());
        y.insert(y.end(), std::make_move_iterator(o.y.begin()), std::make_move_iterator(o.y.end()));
    }
};

template <class V, int METHOD>
class Stepper {
   public:
    using T = typename detail::VecTraits<V>::Scalar;
    using A = detail::Abi<T>;
    using Config = typename A::Config;

    Stats integrate(const Context& ctx) {
        if (METHOD == ODE_B200_DOPRI5 && cfg_.out_type == ODE_B200_OUT_CONTINUOUS)  // dopri5.rs:256-258 panics
            throw std::logic_error("Please use `integrate_with_continuous_output_model` to compute a continuous output model.");
        return run(ctx, nullptr, 0);
    }
    const std::vector<T>& x_out() const { return results_.x; }
    const std::vector<V>& y_out() const { return results_.y; }
    const SolverResult<V>& results() const { return results_; }
    SolverResult<V> into() && { return std::move(results_); }  // impl From<Dopri5<..>> for SolverResult
    Config& config() { return cfg_; }

   protected:
    Stepper(System f, const Config& cfg, V y) : f_(std::move(f)), cfg_(cfg), y_(std::move(y)) {
        if ((size_t)f_.dim != detail::VecTraits<V>::dim(y_)) throw std::invalid_argument("state dimension does not match the system");
    }

    // one trajectory through the host-buffer entry, growing the sample capacity until everything fits
    Stats run(const Context& ctx, ContinuousOutputModel<V>* com, int M) {
        const size_t D = detail::VecTraits<V>::dim(y_);
        std::vector<T> params(f_.params.begin(), f_.params.end());
        int64_t cap = 4096;
        if (METHOD == ODE_B200_RK4 && cfg_.step_size != 0)
            cap = (int64_t)std::max(0.0, std::ceil(((double)cfg_.x_end - (double)cfg_.x0) / (double)cfg_.step_size)) + 2;
        else if (cfg_.out_type == ODE_B200_OUT_DENSE && cfg_.dx > 0)
            cap = (int64_t)std::max(0.0, ((double)cfg_.x_end - (double)cfg_.x0) / (double)cfg_.dx) + 3;
        for (;;) {
            std::vector<T> ox(cap), oy((size_t)cap * D), cb, cs, cc, yf(D);
            if (com) {
                cb.resize(cap);
                cs.resize(cap);
                cc.resize((size_t)cap * M * D);
            }
            T xf = 0, hn = 0, cl = 0;
            uint32_t ne = 0, ac = 0, rj = 0, ns = 0, oc = 0, cn = 0;
            int32_t st = -1;
            typename A::Outputs o{};
            o.y_final = yf.data(); o.x_final = &xf; o.h_next = &hn; o.num_eval = &ne; o.accepted = &ac;
            o.rejected = &rj; o.n_step = &ns; o.status = &st; o.out_cap = cap; o.out_x = ox.data();
            o.out_y = oy.data(); o.out_count = &oc; o.com_cap = com ? cap : 0; o.com_lower = &cl;
            o.com_break = cb.data(); o.com_step = cs.data(); o.com_coef = cc.data(); o.com_count = &cn;
            check(A::integrate(ctx.get(), f_.sys_id, &cfg_, 1, y_.data(), params.data(), 0, &o), "ode_b200_integrate_batch");
            if ((int64_t)oc > cap || (com && (int64_t)cn > cap)) {
                cap = std::max<int64_t>(oc, cn);
                continue;
            }
            results_.x.assign(ox.begin(), ox.begin() + oc);
            results_.y.clear();
            for (uint32_t s = 0; s < oc; ++s) results_.y.push_back(detail::VecTraits<V>::from_flat(&oy[(size_t)s * D], D));
            stats_ = Stats{ne, ac, rj};
            if (com) {
                com->lower_bound = cl;  // set_lower_bound, dopri5.rs:279-281
                for (uint32_t q = 0; q < cn; ++q) {
                    typename ContinuousOutputModel<V>::Interval iv;
                    iv.step_size = cs[q];
                    for (int r = 0; r < M; ++r)
                        iv.coefficients.push_back(detail::VecTraits<V>::from_flat(&cc[((size_t)q * M + r) * D], D));
                    com->breakpoints.push_back(cb[q]);
                    com->intervals.push_back(std::move(iv));
                }
            }
            if (st != ODE_B200_OK && st != ODE_B200_OUTPUT_CAPACITY_EXCEEDED) throw IntegrationError(st, (double)xf, ns);
            return stats_;
        }
    }

    System f_;
    Config cfg_;
    V y_;
    SolverResult<V> results_;
    Stats stats_;
};

template <class V>
class Dopri5V : public Stepper<V, ODE_B200_DOPRI5> {
    using Base = Stepper<V, ODE_B200_DOPRI5>;
    using T = typename Base::T;
    using A = typename Base::A;
    Dopri5V(System f, const typename Base::Config& c, V y) : Base(std::move(f), c, std::move(y)) {}

   public:
    static Dopri5V new_(System f, T x, T x_end, T dx, V y, T rtol, T atol) {  // dopri5.rs:76
        typename Base::Config c;
        check(A::config_new(ODE_B200_DOPRI5, x, x_end, dx, rtol, atol, &c), "config_new");
        return Dopri5V(std::move(f), c, std::move(y));
    }
    static Dopri5V from_param(System f, T x, T x_end, T dx, V y, T rtol, T atol, T safety_factor, T beta, T fac_min,
                              T fac_max, T h_max, T h, uint32_t n_max, uint32_t n_stiff, OutputType out_type) {  // :129
        typename Base::Config c;
        check(A::config_from_param(ODE_B200_DOPRI5, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min, fac_max, h_max, h,
                                   n_max, n_stiff, (int)out_type, &c),
              "config_from_param");
        return Dopri5V(std::move(f), c, std::move(y));
    }
    Stats integrate_with_continuous_output_model(const Context& ctx, ContinuousOutputModel<V>& model) {  // :246
        this->cfg_.out_type = ODE_B200_OUT_CONTINUOUS;  // dopri5.rs:250
        return this->run(ctx, &model, 5);
    }
};

template <class V>
class Dop853V : public Stepper<V, ODE_B200_DOP853> {
    using Base = Stepper<V, ODE_B200_DOP853>;
    using T = typename Base::T;
    using A = typename Base::A;
    Dop853V(System f, const typename Base::Config& c, V y) : Base(std::move(f), c, std::move(y)) {}

   public:
    static Dop853V new_(System f, T x, T x_end, T dx, V y, T rtol, T atol) {  // dop853.rs:76
        typename Base::Config c;
        check(A::config_new(ODE_B200_DOP853, x, x_end, dx, rtol, atol, &c), "config_new");
        return Dop853V(std::move(f), c, std::move(y));
    }
    static Dop853V from_param(System f, T x, T x_end, T dx, V y, T rtol, T atol, T safety_factor, T beta, T fac_min,
                              T fac_max, T h_max, T h, uint32_t n_max, uint32_t n_stiff, OutputType out_type) {  // :133
        typename Base::Config c;
        check(A::config_from_param(ODE_B200_DOP853, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min, fac_max, h_max, h,
                                   n_max, n_stiff, (int)out_type, &c),
              "config_from_param");
        return Dop853V(std::move(f), c, std::move(y));
    }
    // EXTENSION (the crate's Dop853 has no continuous output model): Dense-mode numerics, one interval of the eight
    // rcont vectors per accepted step (ODE_B200_OUT_CONTINUOUS8); evaluate() then returns Dop853's own dense output
    Stats integrate_with_continuous_output_model(const Context& ctx, ContinuousOutputModel<V>& model) {
        static_assert(std::is_same<T, double>::value, "the Continuous8 extension exists for f64 only");
        this->cfg_.out_type = ODE_B200_OUT_CONTINUOUS8;
        return this->run(ctx, &model, 8);
    }
};

template <class V>
class Rk4V : public Stepper<V, ODE_B200_RK4> {
    using Base = Stepper<V, ODE_B200_RK4>;
    using T = typename Base::T;
    Rk4V(System f, const typename Base::Config& c, V y) : Base(std::move(f), c, std::move(y)) {}

   public:
    // Rk4::new(f, x, y, x_end, step_size) -- src/rk4.rs:42
    static Rk4V new_(System f, T x, V y, T x_end, T step_size) {
        typename Base::Config c;
        check(Base::A::config_rk4(x, x_end, step_size, &c), "config_rk4");
        return Rk4V(std::move(f), c, std::move(y));
    }
};

// the crate's spellings: Dopri5<D> = Dopri5<f64, SVector<f64, D>>, Dopri5<D, float>, and the DVector forms
template <int D, class T = double>
using Dopri5 = Dopri5V<State<D, T>>;
template <int D, class T = double>
using Dop853 = Dop853V<State<D, T>>;
template <int D, class T = double>
using Rk4 = Rk4V<State<D, T>>;
using Dopri5Dyn = Dopri5V<DState<double>>;
using Dop853Dyn = Dop853V<DState<double>>;
using Rk4Dyn = Rk4V<DState<double>>;

// ---- the batched calls (new with the ABI) ------------------------------------------------------
// Results of a batched call, structure-of-arrays as in include/ode_b200.h
template <class T>
struct BatchResultT {
    std::vector<T> y_final;  // [dim][n]
    std::vector<T> x_final, h_next;
    std::vector<uint32_t> num_eval, accepted, rejected, n_step;
    std::vector<int32_t> status;
    explicit BatchResultT(size_t dim = 0, size_t n = 0)
        : y_final(dim * n), x_final(n), h_next(n), num_eval(n), accepted(n), rejected(n), n_step(n), status(n, -1) {}
    template <class O>
    void bind(O& o) {
        o.y_final = y_final.data(); o.x_final = x_final.data(); o.h_next = h_next.data(); o.num_eval = num_eval.data();
        o.accepted = accepted.data(); o.rejected = rejected.data(); o.n_step = n_step.data(); o.status = status.data();
    }
};
using BatchResult = BatchResultT<double>;
using BatchResultF32 = BatchResultT<float>;

// n trajectories, one configuration; y0 is [dim][n]; params_per_traj (if given) [n_params][n]. n_gpus > 1 shards the batch
// into contiguous ranges over the node's GPUs (ode_b200_integrate_batch_multi, which makes its own contexts).
inline BatchResult integrate_batch(const Context& ctx, const System& f, const ode_b200_config& cfg, int64_t n, const double* y0,
                                   const double* params_per_traj = nullptr, int n_gpus = 1) {
    BatchResult r((size_t)f.dim, (size_t)n);
    ode_b200_outputs o{};
    r.bind(o);
    const double* pp = params_per_traj ? params_per_traj : f.params.data();
    if (n_gpus > 1)
        check(ode_b200_integrate_batch_multi(f.sys_id, &cfg, n, y0, pp, params_per_traj ? 1 : 0, n_gpus, &o),
              "ode_b200_integrate_batch_multi");
    else
        check(ode_b200_integrate_batch(ctx.get(), f.sys_id, &cfg, n, y0, pp, params_per_traj ? 1 : 0, &o), "ode_b200_integrate_batch");
    return r;
}
inline BatchResultF32 integrate_batch_f32(const Context& ctx, const System& f, const ode_b200_config_f32& cfg, int64_t n,
                                          const float* y0, const float* params_per_traj = nullptr) {
    BatchResultF32 r((size_t)f.dim, (size_t)n);
    ode_b200_outputs_f32 o{};
    r.bind(o);
    std::vector<float> p(f.params.begin(), f.params.end());
    check(ode_b200_integrate_batch_f32(ctx.get(), f.sys_id, &cfg, n, y0, params_per_traj ? params_per_traj : p.data(),
                                       params_per_traj ? 1 : 0, &o),
          "ode_b200_integrate_batch_f32");
    return r;
}
// bouncing_ball.rs's Rk4::<f32> as one call
inline BatchResultF32 rk4_f32_integrate_batch(const Context& ctx, const System& f, float x, float x_end, float step_size, int64_t n,
                                              const float* y0, const float* params_per_traj = nullptr) {
    BatchResultF32 r((size_t)f.dim, (size_t)n);
    ode_b200_outputs_f32 o{};
    r.bind(o);
    std::vector<float> p(f.params.begin(), f.params.end());
    check(ode_b200_rk4_f32_integrate_batch(ctx.get(), f.sys_id, x, x_end, step_size, n, y0,
                                           params_per_traj ? params_per_traj : p.data(), params_per_traj ? 1 : 0, &o),
          "ode_b200_rk4_f32_integrate_batch");
    return r;
}
// device-pointer form: y0, params and every pointer of `out` are device memory of the context's GPU; enqueued on
// `cuda_stream` (nullptr: the default stream) without synchronising
inline void integrate_batch_device(const Context& ctx, const System& f, const ode_b200_config& cfg, int64_t n, const double* d_y0,
                                   const double* d_params, bool params_per_traj, const ode_b200_outputs& out,
                                   void* cuda_stream = nullptr) {
    check(ode_b200_integrate_batch_device(ctx.get(), f.sys_id, &cfg, n, d_y0, d_params, params_per_traj ? 1 : 0, &out, cuda_stream),
          "ode_b200_integrate_batch_device");
}

// ContinuousOutputModel::evaluate for n models (the arrays a Continuous run returned: com_lower [n], com_break /
// com_step [n][cap], com_coef [n][cap][m][dim], com_count [n]) at nq common query points: y [n][nq][dim], found [n][nq]
inline std::pair<std::vector<double>, std::vector<uint8_t>> com_evaluate_batch(
    const Context& ctx, int64_t n, int dim, int m, int64_t cap, const double* com_lower, const double* com_break,
    const double* com_step, const double* com_coef, const uint32_t* com_count, const std::vector<double>& xq) {
    std::vector<double> y((size_t)n * xq.size() * dim);
    std::vector<uint8_t> found((size_t)n * xq.size());
    check(ode_b200_com_evaluate(ctx.get(), n, dim, m, cap, com_lower, com_break, com_step, com_coef, com_count, (int64_t)xq.size(),
                                xq.data(), y.data(), found.data()),
          "ode_b200_com_evaluate");
    return {std::move(y), std::move(found)};
}
// the same with every pointer in device memory (the model where ode_b200_integrate_batch_device left it)
inline void com_evaluate_device(const Context& ctx, int64_t n, int dim, int m, int64_t cap, const double* d_lower,
                                const double* d_break, const double* d_step, const double* d_coef, const uint32_t* d_count,
                                int64_t nq, const double* d_xq, double* d_y, uint8_t* d_found, void* cuda_stream = nullptr) {
    check(ode_b200_com_evaluate_device(ctx.get(), n, dim, m, cap, d_lower, d_break, d_step, d_coef, d_count, nq, d_xq, d_y, d_found,
                                       cuda_stream),
          "ode_b200_com_evaluate_device");
}

}  // namespace ode_b200_host
