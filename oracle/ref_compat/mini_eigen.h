// oracle/ref_compat/mini_eigen.h -- TEST INFRASTRUCTURE ONLY.
// A small eager dense-matrix stand-in for the part of Eigen the reference's IMU code is written against
// (src/preintegration.cpp, include/lvio_fusion/ceres/imu_error.hpp, include/lvio_fusion/utility.h:96-140), so that those
// files compile unchanged without Eigen installed: Matrix<S,R,C[,RowMajor]>, MatrixXd, block<>() / bottomRightCorner<>(),
// the comma initialiser, Map<>, Quaternion<S>, LLT<>, inverse().  Everything is double, sizes are checked at run time.
// [upstream] where the numerics matter: Quaternion * vector uses Eigen's  v + w t + u x t,  t = 2 u x v ; inverse() is a
// partial-pivot LU solve against the identity; LLT is the unblocked lower Cholesky with Eigen's early return on a non-positive pivot.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace Eigen {

enum { ColMajor = 0, RowMajor = 1, Dynamic = -1 };

template <class S, int R, int C, int Opt> class Matrix;
typedef Matrix<double, Dynamic, Dynamic, 0> MatrixXd;
class Block;

// all storage and arithmetic lives here (row-major, run-time sizes)
class Dyn {
public:
    int r = 0, c = 0;
    std::vector<double> a;
    Dyn() {}
    Dyn(int rows, int cols) : r(rows), c(cols), a((size_t)rows * cols, 0.0) {}
    int rows() const { return r; }
    int cols() const { return c; }
    double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
    double& operator()(int i) { return a[(size_t)i]; }
    double operator()(int i) const { return a[(size_t)i]; }
    double& operator[](int i) { return a[(size_t)i]; }
    double operator[](int i) const { return a[(size_t)i]; }
    double& x() { return a[0]; } double& y() { return a[1]; } double& z() { return a[2]; }
    double x() const { return a[0]; } double y() const { return a[1]; } double z() const { return a[2]; }
    double* data() { return a.data(); }
    const double* data() const { return a.data(); }
    void setZero() { for (double& v : a) v = 0.0; }
    void setIdentity() { setZero(); for (int i = 0; i < r && i < c; ++i) (*this)(i, i) = 1.0; }
    double squaredNorm() const { double s = 0; for (double v : a) s += v * v; return s; }
    double norm() const { return std::sqrt(squaredNorm()); }
    void normalize() { const double n = norm(); for (double& v : a) v /= n; }
    inline MatrixXd transpose() const;
    inline MatrixXd inverse() const;
    inline MatrixXd normalized() const;
    inline MatrixXd cross(const Dyn& o) const;
    double dot(const Dyn& o) const { double s = 0; for (size_t i = 0; i < a.size(); ++i) s += a[i] * o.a[i]; return s; }
    bool operator==(const Dyn& o) const { return r == o.r && c == o.c && a == o.a; }
    template <int BR, int BC> inline Block block(int i, int j);
    template <int BR, int BC> inline MatrixXd block(int i, int j) const;
    template <int BR, int BC> inline MatrixXd bottomRightCorner() const;
    Dyn& operator+=(const Dyn& o) { assert(r == o.r && c == o.c); for (size_t i = 0; i < a.size(); ++i) a[i] += o.a[i]; return *this; }
    Dyn& operator-=(const Dyn& o) { assert(r == o.r && c == o.c); for (size_t i = 0; i < a.size(); ++i) a[i] -= o.a[i]; return *this; }
    Dyn& operator*=(double s) { for (double& v : a) v *= s; return *this; }
    Dyn& operator/=(double s) { for (double& v : a) v /= s; return *this; }
};

template <class Derived> class MatrixBase : public Dyn {
public:
    MatrixBase() {}
    MatrixBase(int rows, int cols) : Dyn(rows, cols) {}
};

// comma initialiser:  m << a, b, c, ...;
class CommaInit {
public:
    CommaInit(Dyn& m, double first) : m_(m), k_(0) { put(first); }
    CommaInit& operator,(double v) { put(v); return *this; }
private:
    void put(double v) { assert(k_ < (int)m_.a.size()); m_.a[(size_t)k_++] = v; }
    Dyn& m_;
    int k_;
};

template <class S, int R, int C, int Opt = 0>
class Matrix : public MatrixBase<Matrix<S, R, C, Opt>> {
public:
    typedef S Scalar;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = C, Options = Opt };
    static constexpr bool kFixed2 = R != Dynamic && C != Dynamic && R * C == 2;
    Matrix() : MatrixBase<Matrix>(R == Dynamic ? 0 : R, C == Dynamic ? 0 : C) {}
    // Eigen's rule: on a fixed-size 2-vector two integers are coefficients (Vector2d error(0, 0), backend.cpp:187), else sizes
    Matrix(int rows, int cols) : MatrixBase<Matrix>(kFixed2 ? R : rows, kFixed2 ? C : cols) { if (kFixed2) this->a = {(double)rows, (double)cols}; }
    Matrix(const Dyn& o) : MatrixBase<Matrix>(o.r, o.c) { check(o); this->a = o.a; }                       // NOLINT implicit
    Matrix(double x, double y) : MatrixBase<Matrix>(R, C) { static_assert(R * C == 2, "2-vector"); this->a = {x, y}; }
    Matrix(double x, double y, double z) : MatrixBase<Matrix>(R, C) { static_assert(R * C == 3, "3-vector"); this->a = {x, y, z}; }
    inline Matrix(const Block& b);                                                                         // NOLINT implicit
    Matrix& operator=(const Dyn& o) { check(o); this->r = o.r; this->c = o.c; this->a = o.a; return *this; }
    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(int rows, int cols) { return Matrix(rows, cols); }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    static Matrix UnitX() { Matrix m; m.a[0] = 1.0; return m; }
    static Matrix UnitY() { Matrix m; m.a[1] = 1.0; return m; }
    static Matrix UnitZ() { Matrix m; m.a[2] = 1.0; return m; }
    static Matrix Identity(int rows, int cols) { Matrix m(rows, cols); m.setIdentity(); return m; }
    CommaInit operator<<(double first) { return CommaInit(*this, first); }
private:
    void check(const Dyn& o) const { if (!((R == Dynamic || R == o.r) && (C == Dynamic || C == o.c))) { assert(!"matrix size mismatch"); std::abort(); } }
};
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 3, 3> Matrix3d;
class VectorXd : public Matrix<double, Dynamic, 1> {           // environment.cpp:124: VectorXd v(3)
public:
    explicit VectorXd(int n) : Matrix<double, Dynamic, 1>(n, 1) {}
    VectorXd(const Dyn& o) : Matrix<double, Dynamic, 1>(o) {}                                               // NOLINT implicit
};

inline MatrixXd operator+(const Dyn& x, const Dyn& y) { assert(x.r == y.r && x.c == y.c); MatrixXd o(x.r, x.c); for (size_t i = 0; i < x.a.size(); ++i) o.a[i] = x.a[i] + y.a[i]; return o; }
inline MatrixXd operator-(const Dyn& x, const Dyn& y) { assert(x.r == y.r && x.c == y.c); MatrixXd o(x.r, x.c); for (size_t i = 0; i < x.a.size(); ++i) o.a[i] = x.a[i] - y.a[i]; return o; }
inline MatrixXd operator-(const Dyn& x) { MatrixXd o(x.r, x.c); for (size_t i = 0; i < x.a.size(); ++i) o.a[i] = -x.a[i]; return o; }
inline MatrixXd operator*(const Dyn& x, double s) { MatrixXd o(x.r, x.c); for (size_t i = 0; i < x.a.size(); ++i) o.a[i] = x.a[i] * s; return o; }
inline MatrixXd operator*(double s, const Dyn& x) { MatrixXd o(x.r, x.c); for (size_t i = 0; i < x.a.size(); ++i) o.a[i] = s * x.a[i]; return o; }
inline MatrixXd operator/(const Dyn& x, double s) { MatrixXd o(x.r, x.c); for (size_t i = 0; i < x.a.size(); ++i) o.a[i] = x.a[i] / s; return o; }
inline MatrixXd operator*(const Dyn& x, const Dyn& y) {
    assert(x.c == y.r);
    MatrixXd o(x.r, y.c);
    for (int i = 0; i < x.r; ++i) for (int j = 0; j < y.c; ++j) { double s = 0; for (int k = 0; k < x.c; ++k) s += x(i, k) * y(k, j); o(i, j) = s; }
    return o;
}
inline MatrixXd Dyn::transpose() const { MatrixXd o(c, r); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) o(j, i) = (*this)(i, j); return o; }
inline MatrixXd Dyn::normalized() const { return *this / norm(); }
inline MatrixXd Dyn::cross(const Dyn& o) const { MatrixXd m(3, 1); m.a = {a[1] * o.a[2] - a[2] * o.a[1], a[2] * o.a[0] - a[0] * o.a[2], a[0] * o.a[1] - a[1] * o.a[0]}; return m; }
inline MatrixXd Dyn::inverse() const {          // partial-pivot LU, solve against the identity
    assert(r == c);
    const int n = r;
    std::vector<double> lu = a; std::vector<int> piv(n);
    for (int i = 0; i < n; ++i) piv[i] = i;
    for (int k = 0; k < n; ++k) {
        int best = k; double bv = std::fabs(lu[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) if (std::fabs(lu[(size_t)i * n + k]) > bv) { bv = std::fabs(lu[(size_t)i * n + k]); best = i; }
        if (best != k) { for (int j = 0; j < n; ++j) std::swap(lu[(size_t)k * n + j], lu[(size_t)best * n + j]); std::swap(piv[k], piv[best]); }
        for (int i = k + 1; i < n; ++i) { lu[(size_t)i * n + k] /= lu[(size_t)k * n + k]; const double f = lu[(size_t)i * n + k]; for (int j = k + 1; j < n; ++j) lu[(size_t)i * n + j] -= f * lu[(size_t)k * n + j]; }
    }
    MatrixXd inv(n, n);
    std::vector<double> y(n);
    for (int col = 0; col < n; ++col) {
        for (int i = 0; i < n; ++i) { double s = piv[i] == col ? 1.0 : 0.0; for (int k = 0; k < i; ++k) s -= lu[(size_t)i * n + k] * y[k]; y[i] = s; }
        for (int i = n - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < n; ++k) s -= lu[(size_t)i * n + k] * inv(k, col); inv(i, col) = s / lu[(size_t)i * n + i]; }
    }
    return inv;
}

// writable view (base pointer + strides): blocks of matrices and of Map<>
class Block {
public:
    Block(double* base, int rs, int cs, int rows, int cols) : p_(base), rs_(rs), cs_(cs), r_(rows), c_(cols) {}
    double& operator()(int i, int j) { return p_[(size_t)i * rs_ + (size_t)j * cs_]; }
    double operator()(int i, int j) const { return p_[(size_t)i * rs_ + (size_t)j * cs_]; }
    int rows() const { return r_; }
    int cols() const { return c_; }
    Block& operator=(const Dyn& m) { assert(m.r == r_ && m.c == c_); for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) (*this)(i, j) = m(i, j); return *this; }
    Block& operator=(const Block& o) { MatrixXd t(o.r_, o.c_); for (int i = 0; i < o.r_; ++i) for (int j = 0; j < o.c_; ++j) t(i, j) = o(i, j); return *this = static_cast<const Dyn&>(t); }
    inline MatrixXd eval() const;
private:
    double* p_; int rs_, cs_, r_, c_;
};
inline MatrixXd Block::eval() const { MatrixXd t(r_, c_); for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) t(i, j) = (*this)(i, j); return t; }
inline MatrixXd operator*(const Dyn& x, const Block& y) { return x * static_cast<const Dyn&>(y.eval()); }
inline MatrixXd operator*(const Block& x, const Dyn& y) { return static_cast<const Dyn&>(x.eval()) * y; }
inline MatrixXd operator+(const Dyn& x, const Block& y) { return x + static_cast<const Dyn&>(y.eval()); }
inline MatrixXd operator-(const Dyn& x, const Block& y) { return x - static_cast<const Dyn&>(y.eval()); }
template <class S, int R, int C, int Opt> inline Matrix<S, R, C, Opt>::Matrix(const Block& b) : MatrixBase<Matrix>(b.rows(), b.cols()) { const MatrixXd t = b.eval(); check(t); this->a = t.a; }
template <int BR, int BC> inline Block Dyn::block(int i, int j) { assert(i + BR <= r && j + BC <= c); return Block(a.data() + (size_t)i * c + j, c, 1, BR, BC); }
template <int BR, int BC> inline MatrixXd Dyn::block(int i, int j) const { MatrixXd t(BR, BC); for (int u = 0; u < BR; ++u) for (int v = 0; v < BC; ++v) t(u, v) = (*this)(i + u, j + v); return t; }
template <int BR, int BC> inline MatrixXd Dyn::bottomRightCorner() const { return block<BR, BC>(r - BR, c - BC); }

// Map<Matrix<double, R, C[, RowMajor]>> over caller memory
template <class M> class Map {
public:
    enum { R = M::RowsAtCompileTime, C = M::ColsAtCompileTime, RM = (M::Options & RowMajor) ? 1 : 0 };
    explicit Map(double* p) : p_(p) {}
    double& operator()(int i, int j) { return RM ? p_[(size_t)i * C + j] : p_[(size_t)j * R + i]; }
    double operator()(int i, int j) const { return RM ? p_[(size_t)i * C + j] : p_[(size_t)j * R + i]; }
    void setZero() { for (int i = 0; i < R * C; ++i) p_[i] = 0.0; }
    template <int BR, int BC> Block block(int i, int j) { return Block(&(*this)(i, j), RM ? C : 1, RM ? 1 : R, BR, BC); }
    MatrixXd eval() const { MatrixXd t(R, C); for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) t(i, j) = (*this)(i, j); return t; }
    Map& operator=(const Dyn& m) { assert(m.r == R && m.c == C); for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) (*this)(i, j) = m(i, j); return *this; }
private:
    double* p_;
};
template <class M> inline MatrixXd operator*(const Dyn& x, const Map<M>& y) { return x * static_cast<const Dyn&>(y.eval()); }

// LLT<M>(A).matrixL().transpose().  Eigen 3.3 Cholesky/LLT.h, llt_inplace<Scalar, Lower>::unblocked (what blocked() calls
// below 32 rows), restated INCLUDING its failure path: the factorisation runs in place on a copy of A, reads the lower
// triangle only, and on the first pivot x <= 0 returns at once -- LLT::info() turns NumericalIssue, which the reference
// never looks at (imu_error.hpp:32,150,257) -- so matrixL() then shows the finished columns followed by the untouched lower
// triangle of A.  ImuInitError reaches that path with the reference's own settings (priors 1e4 / 1e2 written over the
// bias blocks of cov^-1, initializer.cpp:62, make the matrix indefinite), so it is part of the behaviour to reproduce.
template <class M> class LLT {
public:
    explicit LLT(const Dyn& A) : L_(A.r, A.c) {
        const int n = A.r;
        for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) L_(i, j) = A(i, j);
        for (int k = 0; k < n; ++k) {
            double x = L_(k, k);
            for (int j = 0; j < k; ++j) x -= L_(k, j) * L_(k, j);
            if (x <= 0.0) break;
            L_(k, k) = x = std::sqrt(x);
            for (int i = k + 1; i < n; ++i) { double s = L_(i, k); for (int j = 0; j < k; ++j) s -= L_(i, j) * L_(k, j); L_(i, k) = s / x; }
        }
    }
    const MatrixXd& matrixL() const { return L_; }
private:
    MatrixXd L_;
};

// ---- quaternions (storage x y z w like Eigen; constructor order w x y z)
template <class Derived> class QuaternionBase {
public:
    const Derived& derived() const { return *static_cast<const Derived*>(this); }
    double w() const { return derived().q[3]; } double x() const { return derived().q[0]; } double y() const { return derived().q[1]; } double z() const { return derived().q[2]; }
    Vector3d vec() const { return Vector3d(x(), y(), z()); }
};
template <class S> class Quaternion : public QuaternionBase<Quaternion<S>> {
public:
    typedef S Scalar;
    double q[4];
    Quaternion() : q{0, 0, 0, 1} {}
    Quaternion(double w, double x, double y, double z) : q{x, y, z, w} {}
    explicit Quaternion(const Dyn& m) {                  // from a rotation matrix (Eigen's quaternionbase_assign_impl)
        const double t = m(0, 0) + m(1, 1) + m(2, 2);
        if (t > 0) { double s = std::sqrt(t + 1.0); q[3] = 0.5 * s; s = 0.5 / s; q[0] = (m(2, 1) - m(1, 2)) * s; q[1] = (m(0, 2) - m(2, 0)) * s; q[2] = (m(1, 0) - m(0, 1)) * s; }
        else {
            int i = 0; if (m(1, 1) > m(0, 0)) i = 1; if (m(2, 2) > m(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            double s = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
            q[i] = 0.5 * s; s = 0.5 / s; q[3] = (m(k, j) - m(j, k)) * s; q[j] = (m(j, i) + m(i, j)) * s; q[k] = (m(k, i) + m(i, k)) * s;
        }
    }
    using QuaternionBase<Quaternion<S>>::w; using QuaternionBase<Quaternion<S>>::x; using QuaternionBase<Quaternion<S>>::y; using QuaternionBase<Quaternion<S>>::z;
    double& w() { return q[3]; } double& x() { return q[0]; } double& y() { return q[1]; } double& z() { return q[2]; }
    struct Coeffs { double* p; double* data() { return p; } const double* data() const { return p; } double& operator[](int i) { return p[i]; } };      // (x, y, z, w), Eigen's storage order
    Coeffs coeffs() { return Coeffs{q}; }
    Coeffs coeffs() const { return Coeffs{const_cast<double*>(q)}; }
    static Quaternion Identity() { return Quaternion(); }
    void setIdentity() { q[0] = q[1] = q[2] = 0; q[3] = 1; }
    double squaredNorm() const { return q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]; }
    double norm() const { return std::sqrt(squaredNorm()); }
    void normalize() { const double n = norm(); for (double& v : q) v /= n; }
    Quaternion normalized() const { Quaternion o = *this; o.normalize(); return o; }
    Quaternion slerp(double t, const Quaternion& o) const {          // Eigen/src/Geometry/Quaternion.h: QuaternionBase::slerp
        const double one = 1.0 - 2.220446049250313e-16;
        const double d = q[0] * o.q[0] + q[1] * o.q[1] + q[2] * o.q[2] + q[3] * o.q[3], ad = std::fabs(d);
        double s0, s1;
        if (ad >= one) { s0 = 1.0 - t; s1 = t; }
        else { const double th = std::acos(ad), st = std::sin(th); s0 = std::sin((1.0 - t) * th) / st; s1 = std::sin(t * th) / st; }
        if (d < 0) s1 = -s1;
        return Quaternion(s0 * q[3] + s1 * o.q[3], s0 * q[0] + s1 * o.q[0], s0 * q[1] + s1 * o.q[1], s0 * q[2] + s1 * o.q[2]);
    }
    Quaternion conjugate() const { return Quaternion(q[3], -q[0], -q[1], -q[2]); }
    Quaternion inverse() const { const double n2 = squaredNorm(); Quaternion c = conjugate(); for (double& v : c.q) v /= n2; return c; }
    Quaternion operator*(const Quaternion& b) const {
        const double aw = q[3], ax = q[0], ay = q[1], az = q[2], bw = b.q[3], bx = b.q[0], by = b.q[1], bz = b.q[2];
        return Quaternion(aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx);
    }
    Matrix<double, 3, 3> operator*(const Matrix<double, 3, 3>& m) const { return Matrix<double, 3, 3>(toRotationMatrix() * m); }      // RotationBase * matrix (map.cpp:112)
    Vector3d operator*(const Dyn& v) const {            // Eigen's _transformVector
        const Vector3d u(q[0], q[1], q[2]);
        const Vector3d t = 2.0 * u.cross(v);
        return Vector3d(v + q[3] * t + u.cross(t));
    }
    Matrix3d toRotationMatrix() const {                 // Eigen's QuaternionBase::toRotationMatrix
        const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
        const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0], tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
        Matrix3d m;
        m << 1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy);
        return m;
    }
};
typedef Quaternion<double> Quaterniond;

}  // namespace Eigen
