// oracle/dual.h -- forward-mode dual numbers for the CPU oracle.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path;
// it is the checker the CUDA kernels are compared against (tests/, smoke(), and
// bench.py's cpu_baseline leg).
//
// Stands in for ceres::Jet<double,N> (un-vendored third party, Ceres 2.0-2.1 inferred
// from /root/reference/src/lvio_fusion/include/lvio_fusion/adapt/problem.h:44): an
// AutoDiffCostFunction returns the exact derivative, so any correct forward-mode AD
// reproduces it to rounding.  "parity unpinned": the reference ships no golden vectors.
#pragma once
#include <cmath>

namespace oracle {

template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0.0) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
    Dual(double s) : v(s) { for (int i = 0; i < N; ++i) d[i] = 0.0; }  // NOLINT implicit
    static Dual seed(double s, int k) { Dual r(s); r.d[k] = 1.0; return r; }
};

template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) {
    Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; const double inv = 1.0 / b.v; r.v = a.v * inv;
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv; return r; }

template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double a, const Dual<N>& b) { return b + a; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double a, const Dual<N>& b) { return (-b) + a; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) {
    Dual<N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <int N> inline Dual<N> operator*(double a, const Dual<N>& b) { return b * a; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { return a * (1.0 / b); }
template <int N> inline Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }
template <int N> inline Dual<N>& operator+=(Dual<N>& a, const Dual<N>& b) { a = a + b; return a; }
template <int N> inline Dual<N>& operator-=(Dual<N>& a, const Dual<N>& b) { a = a - b; return a; }
template <int N> inline Dual<N>& operator*=(Dual<N>& a, const Dual<N>& b) { a = a * b; return a; }

template <int N> inline Dual<N> sqrt(const Dual<N>& a) {
    Dual<N> r; r.v = std::sqrt(a.v); const double k = 0.5 / r.v;
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k; return r; }
template <int N> inline Dual<N> sin(const Dual<N>& a) {
    Dual<N> r; r.v = std::sin(a.v); const double k = std::cos(a.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k; return r; }
template <int N> inline Dual<N> cos(const Dual<N>& a) {
    Dual<N> r; r.v = std::cos(a.v); const double k = -std::sin(a.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k; return r; }
template <int N> inline Dual<N> asin(const Dual<N>& a) {
    Dual<N> r; r.v = std::asin(a.v); const double k = 1.0 / std::sqrt(1.0 - a.v * a.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k; return r; }
template <int N> inline Dual<N> atan2(const Dual<N>& y, const Dual<N>& x) {
    Dual<N> r; r.v = std::atan2(y.v, x.v); const double k = 1.0 / (x.v * x.v + y.v * y.v);
    for (int i = 0; i < N; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) * k; return r; }

// scalar overloads so templated code can say sqrt(T) etc. through ADL-free calls
inline double sqrt(double a) { return std::sqrt(a); }
inline double sin(double a) { return std::sin(a); }
inline double cos(double a) { return std::cos(a); }
inline double asin(double a) { return std::asin(a); }
inline double atan2(double y, double x) { return std::atan2(y, x); }
inline float sqrt(float a) { return std::sqrt(a); }

inline double value_of(double a) { return a; }
template <int N> inline double value_of(const Dual<N>& a) { return a.v; }

}  // namespace oracle
