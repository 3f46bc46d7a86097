// include/lvio_b200/ceres_autodiff.h -- the header-only pieces of Ceres the reference's *off-path* functors are written
// against: ceres::Jet, ceres::AutoDiffCostFunction and the three ceres/rotation.h templates its
// helpers call (/root/reference/src/lvio_fusion/include/lvio_fusion/ceres/base.hpp:30,63,82 -> QuaternionRotatePoint,
// QuaternionProduct, DotProduct; navsat_error.hpp, pose_error.hpp create AutoDiffCostFunction<F, kRes, Ns...>).
//
// These serve the small host-side solves only (navsat alignment, section pose graph, relocation: a handful of unknowns,
// see host_solver.h).  The hot-path factor classes (factors.h) never go through them: they are device records.
// [upstream] Forward-mode dual numbers give the exact derivative, so any correct implementation reproduces Ceres' Jacobians to
// rounding; UnitQuaternionRotatePoint uses the "uv" form of Ceres 2.x.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <utility>
#include <vector>
#include "ceres_shim.h"

namespace ceres {

template <typename T, int N>
struct Jet {
    T a;          // value
    T v[N];       // derivatives
    Jet() : a(T(0)) { for (int i = 0; i < N; ++i) v[i] = T(0); }
    Jet(const T& s) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(0); }   // NOLINT implicit, like ceres::Jet
    Jet(const T& s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(0); v[k] = T(1); }
};
#define LVB_JET template <typename T, int N> inline
LVB_JET Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
LVB_JET Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
LVB_JET Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
LVB_JET Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
LVB_JET Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
LVB_JET Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
    Jet<T, N> h; const T inv = T(1) / g.a; h.a = f.a * inv;
    for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - h.a * g.v[i]) * inv;
    return h;
}
LVB_JET Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a += s; return h; }
LVB_JET Jet<T, N> operator+(T s, const Jet<T, N>& f) { return f + s; }
LVB_JET Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a -= s; return h; }
LVB_JET Jet<T, N> operator-(T s, const Jet<T, N>& f) { return (-f) + s; }
LVB_JET Jet<T, N> operator*(const Jet<T, N>& f, T s) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
LVB_JET Jet<T, N> operator*(T s, const Jet<T, N>& f) { return f * s; }
LVB_JET Jet<T, N> operator/(const Jet<T, N>& f, T s) { return f * (T(1) / s); }
LVB_JET Jet<T, N> operator/(T s, const Jet<T, N>& f) { return Jet<T, N>(s) / f; }
LVB_JET Jet<T, N>& operator+=(Jet<T, N>& f, const Jet<T, N>& g) { f = f + g; return f; }
LVB_JET Jet<T, N>& operator-=(Jet<T, N>& f, const Jet<T, N>& g) { f = f - g; return f; }
LVB_JET Jet<T, N>& operator*=(Jet<T, N>& f, const Jet<T, N>& g) { f = f * g; return f; }
LVB_JET Jet<T, N>& operator/=(Jet<T, N>& f, const Jet<T, N>& g) { f = f / g; return f; }
LVB_JET Jet<T, N>& operator+=(Jet<T, N>& f, T s) { f.a += s; return f; }
LVB_JET Jet<T, N>& operator-=(Jet<T, N>& f, T s) { f.a -= s; return f; }
LVB_JET Jet<T, N>& operator*=(Jet<T, N>& f, T s) { f = f * s; return f; }
LVB_JET Jet<T, N>& operator/=(Jet<T, N>& f, T s) { f = f / s; return f; }
LVB_JET bool operator<(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a < g.a; }
LVB_JET bool operator<=(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a <= g.a; }
LVB_JET bool operator>(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a > g.a; }
LVB_JET bool operator>=(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a >= g.a; }
LVB_JET bool operator==(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a == g.a; }
LVB_JET bool operator!=(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a != g.a; }
LVB_JET bool operator<(const Jet<T, N>& f, T s) { return f.a < s; }
LVB_JET bool operator>(const Jet<T, N>& f, T s) { return f.a > s; }
LVB_JET bool operator<=(const Jet<T, N>& f, T s) { return f.a <= s; }
LVB_JET bool operator>=(const Jet<T, N>& f, T s) { return f.a >= s; }
LVB_JET Jet<T, N> chain(const Jet<T, N>& f, T value, T slope) { Jet<T, N> h; h.a = value; for (int i = 0; i < N; ++i) h.v[i] = slope * f.v[i]; return h; }
LVB_JET Jet<T, N> sqrt(const Jet<T, N>& f) { const T r = std::sqrt(f.a); return chain(f, r, T(1) / (T(2) * r)); }
LVB_JET Jet<T, N> sin(const Jet<T, N>& f) { return chain(f, std::sin(f.a), std::cos(f.a)); }
LVB_JET Jet<T, N> cos(const Jet<T, N>& f) { return chain(f, std::cos(f.a), -std::sin(f.a)); }
LVB_JET Jet<T, N> tan(const Jet<T, N>& f) { const T t = std::tan(f.a); return chain(f, t, T(1) + t * t); }
LVB_JET Jet<T, N> asin(const Jet<T, N>& f) { return chain(f, std::asin(f.a), T(1) / std::sqrt(T(1) - f.a * f.a)); }
LVB_JET Jet<T, N> acos(const Jet<T, N>& f) { return chain(f, std::acos(f.a), -T(1) / std::sqrt(T(1) - f.a * f.a)); }
LVB_JET Jet<T, N> atan(const Jet<T, N>& f) { return chain(f, std::atan(f.a), T(1) / (T(1) + f.a * f.a)); }
LVB_JET Jet<T, N> exp(const Jet<T, N>& f) { const T e = std::exp(f.a); return chain(f, e, e); }
LVB_JET Jet<T, N> log(const Jet<T, N>& f) { return chain(f, std::log(f.a), T(1) / f.a); }
LVB_JET Jet<T, N> abs(const Jet<T, N>& f) { return f.a < T(0) ? -f : f; }
LVB_JET Jet<T, N> pow(const Jet<T, N>& f, T p) { const T r = std::pow(f.a, p); return chain(f, r, p * std::pow(f.a, p - T(1))); }
LVB_JET Jet<T, N> atan2(const Jet<T, N>& y, const Jet<T, N>& x) {
    Jet<T, N> h; h.a = std::atan2(y.a, x.a);
    const T k = T(1) / (x.a * x.a + y.a * y.a);
    for (int i = 0; i < N; ++i) h.v[i] = (x.a * y.v[i] - y.a * x.v[i]) * k;
    return h;
}
LVB_JET bool isfinite(const Jet<T, N>& f) { if (!std::isfinite(f.a)) return false; for (int i = 0; i < N; ++i) if (!std::isfinite(f.v[i])) return false; return true; }
#undef LVB_JET
// the scalar overloads generic functors find through `using` / ADL-free calls inside namespace ceres
using std::abs; using std::acos; using std::asin; using std::atan; using std::atan2; using std::cos; using std::exp; using std::log; using std::pow; using std::sin; using std::sqrt; using std::tan;

// ---- ceres/rotation.h subset (w-first quaternions) -------------------------------------------------------------------
template <typename T> inline T DotProduct(const T x[3], const T y[3]) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }
template <typename T> inline void CrossProduct(const T x[3], const T y[3], T out[3]) {
    out[0] = x[1] * y[2] - x[2] * y[1]; out[1] = x[2] * y[0] - x[0] * y[2]; out[2] = x[0] * y[1] - x[1] * y[0];
}
template <typename T> inline void QuaternionProduct(const T z[4], const T w[4], T zw[4]) {
    zw[0] = z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3];
    zw[1] = z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2];
    zw[2] = z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1];
    zw[3] = z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0];
}
template <typename T> inline void UnitQuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
    T uv0 = q[2] * pt[2] - q[3] * pt[1], uv1 = q[3] * pt[0] - q[1] * pt[2], uv2 = q[1] * pt[1] - q[2] * pt[0];
    uv0 += uv0; uv1 += uv1; uv2 += uv2;
    result[0] = pt[0] + q[0] * uv0; result[1] = pt[1] + q[0] * uv1; result[2] = pt[2] + q[0] * uv2;
    result[0] += q[2] * uv2 - q[3] * uv1; result[1] += q[3] * uv0 - q[1] * uv2; result[2] += q[1] * uv1 - q[2] * uv0;
}
template <typename T> inline void QuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
    const T scale = T(1) / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const T unit[4] = {scale * q[0], scale * q[1], scale * q[2], scale * q[3]};
    UnitQuaternionRotatePoint(unit, pt, result);
}

// ---- AutoDiffCostFunction<Functor, kNumResiduals, N0, N1, ...> (Ceres 2.x variadic form) -----------------------------
namespace internal {
template <int... Ns> struct Sum;
template <> struct Sum<> { static constexpr int value = 0; };
template <int N, int... Ns> struct Sum<N, Ns...> { static constexpr int value = N + Sum<Ns...>::value; };
}  // namespace internal

template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public SizedCostFunction<kNumResiduals, Ns...> {
public:
    explicit AutoDiffCostFunction(Functor* functor, Ownership ownership = TAKE_OWNERSHIP) : functor_(functor), ownership_(ownership) {
        static_assert(kNumResiduals > 0, "dynamic residual counts are not used by the reference");
    }
    ~AutoDiffCostFunction() override { if (ownership_ == TAKE_OWNERSHIP) delete functor_; }
    bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
        constexpr int kBlocks = sizeof...(Ns);
        if (!jacobians) return call<double>(parameters, residuals, std::make_index_sequence<kBlocks>());
        constexpr int kTotal = internal::Sum<Ns...>::value;
        typedef Jet<double, kTotal> J;
        const int sizes[kBlocks] = {Ns...};
        std::vector<J> x(kTotal);
        const J* ptrs[kBlocks];
        int off = 0;
        for (int b = 0; b < kBlocks; ++b) {
            ptrs[b] = x.data() + off;
            for (int k = 0; k < sizes[b]; ++k) x[off + k] = J(parameters[b][k], off + k);
            off += sizes[b];
        }
        J out[kNumResiduals];
        if (!call<J>(ptrs, out, std::make_index_sequence<kBlocks>())) return false;
        off = 0;
        for (int b = 0; b < kBlocks; ++b) {
            if (jacobians[b]) for (int r = 0; r < kNumResiduals; ++r) for (int k = 0; k < sizes[b]; ++k) jacobians[b][r * sizes[b] + k] = out[r].v[off + k];
            off += sizes[b];
        }
        for (int r = 0; r < kNumResiduals; ++r) residuals[r] = out[r].a;
        return true;
    }
private:
    template <typename T, size_t... I>
    bool call(T const* const* p, T* residuals, std::index_sequence<I...>) const { return (*functor_)(p[I]..., residuals); }
    Functor* functor_;
    Ownership ownership_;
};

// ---- NumericDiffCostFunction<Functor, method, kNumResiduals, N0, N1, ...> ------------------------------------------------
// Named once by the reference: ImuInitGError (imu_error.hpp:263-266, FORWARD), the gravity-direction factor of
// imu::InertialOptimization.  [upstream, restated from Ceres' internal/numeric_diff.h; parity unpinned]: one-sided difference
// per coordinate with step  h_j = max(sqrt(DBL_EPSILON), |x_j| * 1e-6)  (NumericDiffOptions::relative_step_size), CENTRAL uses
// the two-sided quotient with the same step.  RIDDERS is not provided.
enum NumericDiffMethodType { CENTRAL, FORWARD, RIDDERS };

template <typename Functor, NumericDiffMethodType kMethod, int kNumResiduals, int... Ns>
class NumericDiffCostFunction : public SizedCostFunction<kNumResiduals, Ns...> {
public:
    explicit NumericDiffCostFunction(Functor* functor, Ownership ownership = TAKE_OWNERSHIP) : functor_(functor), ownership_(ownership) {
        static_assert(kMethod != RIDDERS, "Ridders' method is not provided");
        static_assert(kNumResiduals > 0, "dynamic residual counts are not used by the reference");
    }
    ~NumericDiffCostFunction() override { if (ownership_ == TAKE_OWNERSHIP) delete functor_; }
    bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
        constexpr int kBlocks = sizeof...(Ns);
        if (!call(parameters, residuals, std::make_index_sequence<kBlocks>())) return false;
        if (!jacobians) return true;
        const int sizes[kBlocks] = {Ns...};
        // a private copy of the parameters: one coordinate is perturbed at a time
        std::vector<std::vector<double>> copy(kBlocks);
        const double* ptrs[kBlocks];
        for (int b = 0; b < kBlocks; ++b) { copy[b].assign(parameters[b], parameters[b] + sizes[b]); ptrs[b] = copy[b].data(); }
        const double min_step = std::sqrt(std::numeric_limits<double>::epsilon());
        double plus[kNumResiduals], minus[kNumResiduals];
        for (int b = 0; b < kBlocks; ++b) {
            if (!jacobians[b]) continue;
            for (int k = 0; k < sizes[b]; ++k) {
                const double x = parameters[b][k], h = std::max(min_step, std::fabs(x) * 1e-6);
                copy[b][k] = x + h;
                if (!call(ptrs, plus, std::make_index_sequence<kBlocks>())) return false;
                if (kMethod == CENTRAL) {
                    copy[b][k] = x - h;
                    if (!call(ptrs, minus, std::make_index_sequence<kBlocks>())) return false;
                    for (int r = 0; r < kNumResiduals; ++r) jacobians[b][r * sizes[b] + k] = (plus[r] - minus[r]) / (2 * h);
                } else {
                    for (int r = 0; r < kNumResiduals; ++r) jacobians[b][r * sizes[b] + k] = (plus[r] - residuals[r]) / h;
                }
                copy[b][k] = x;
            }
        }
        return true;
    }
private:
    template <size_t... I>
    bool call(double const* const* p, double* residuals, std::index_sequence<I...>) const { return (*functor_)(p[I]..., residuals); }
    Functor* functor_;
    Ownership ownership_;
};

}  // namespace ceres
