// oracle/geometry.h -- rigid-body helpers of the CPU oracle (TEST INFRASTRUCTURE ONLY).
//
// Restates, in value-returning struct form, the helper set the reference's factors are
// written against:
//   /root/reference/src/lvio_fusion/include/lvio_fusion/ceres/base.hpp:10-157
// plus the three ceres/rotation.h templates those helpers call (Ceres is NOT vendored in
// the reference and no version is pinned; behaviour below is "upstream, unverified
// in-container" and marked [upstream]).
//
// Storage convention everywhere: a pose is Sophus::SE3d::data() = [qx qy qz qw tx ty tz]
// (base.hpp:27-38 reads e_q[3] as w and se3+4 as the translation).
#pragma once
#include "dual.h"

namespace oracle {

template <class T> struct Vec3 {
    T x, y, z;
    Vec3() : x(T(0)), y(T(0)), z(T(0)) {}
    Vec3(T a, T b, T c) : x(a), y(b), z(c) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> inline Vec3<T> operator+(const Vec3<T>& a, const Vec3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }  // base.hpp:19-24
template <class T> inline Vec3<T> operator-(const Vec3<T>& a, const Vec3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }  // base.hpp:11-16
template <class T> inline Vec3<T> operator-(const Vec3<T>& a) { return {-a.x, -a.y, -a.z}; }
template <class T, class S> inline Vec3<T> operator*(const Vec3<T>& a, const S& s) { return {a.x * s, a.y * s, a.z * s}; }
// [upstream] ceres::DotProduct: x0*y0 + x1*y1 + x2*y2, left to right.
template <class T> inline T dot(const Vec3<T>& a, const Vec3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> inline Vec3<T> cross(const Vec3<T>& a, const Vec3<T>& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Quaternion in Eigen storage order (x, y, z, w).
template <class T> struct Quat {
    T x, y, z, w;
    Quat() : x(T(0)), y(T(0)), z(T(0)), w(T(1)) {}
    Quat(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {}
};

template <class T> struct Rigid {  // T_parent_child
    Quat<T> q;
    Vec3<T> t;
};

template <class T, class S> inline Rigid<T> load_rigid(const S* p) {
    Rigid<T> r;
    r.q = Quat<T>(T(p[0]), T(p[1]), T(p[2]), T(p[3]));
    r.t = Vec3<T>(T(p[4]), T(p[5]), T(p[6]));
    return r;
}
template <class T> inline void store_rigid(const Rigid<T>& r, T* p) {
    p[0] = r.q.x; p[1] = r.q.y; p[2] = r.q.z; p[3] = r.q.w; p[4] = r.t.x; p[5] = r.t.y; p[6] = r.t.z;
}

// [upstream] ceres::UnitQuaternionRotatePoint, Ceres 2.x "uv" form (the 1.x releases used
// a t2..t9 product form; same mathematics, different float32 rounding -- unpinned).
template <class T> inline Vec3<T> rotate_unit(const Quat<T>& u, const Vec3<T>& p) {
    T uv0 = u.y * p.z - u.z * p.y;
    T uv1 = u.z * p.x - u.x * p.z;
    T uv2 = u.x * p.y - u.y * p.x;
    uv0 = uv0 + uv0;
    uv1 = uv1 + uv1;
    uv2 = uv2 + uv2;
    Vec3<T> r(p.x + u.w * uv0, p.y + u.w * uv1, p.z + u.w * uv2);
    r.x = r.x + (u.y * uv2 - u.z * uv1);
    r.y = r.y + (u.z * uv0 - u.x * uv2);
    r.z = r.z + (u.x * uv1 - u.y * uv0);
    return r;
}

// base.hpp:26-31 EigenQuaternionRotatePoint -> [upstream] ceres::QuaternionRotatePoint,
// which first rescales q by 1/sqrt(w^2+x^2+y^2+z^2) -- autodiff therefore "sees" the
// normalisation and the 2x4 quaternion Jacobian is tangential to the unit sphere.
template <class T> inline Vec3<T> rotate(const Quat<T>& q, const Vec3<T>& p) {
    const T scale = T(1) / sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    const Quat<T> u(scale * q.x, scale * q.y, scale * q.z, scale * q.w);
    return rotate_unit(u, p);
}

// base.hpp:33-38 SE3TransformPoint: R(q) p + t
template <class T> inline Vec3<T> apply(const Rigid<T>& a, const Vec3<T>& p) { return rotate(a.q, p) + a.t; }

// base.hpp:40-47 EigenQuaternionInverse: conjugate, no normalisation
template <class T> inline Quat<T> conj(const Quat<T>& q) { return Quat<T>(-q.x, -q.y, -q.z, q.w); }

// base.hpp:49-55 SE3Inverse: (conj q, R(conj q)(-t))
template <class T> inline Rigid<T> inverse(const Rigid<T>& a) {
    Rigid<T> r;
    r.q = conj(a.q);
    r.t = rotate(r.q, -a.t);
    return r;
}

// base.hpp:57-68 EigenQuaternionProduct -> [upstream] ceres::QuaternionProduct (Hamilton, w first)
template <class T> inline Quat<T> mul(const Quat<T>& z, const Quat<T>& w) {
    Quat<T> r;
    r.w = z.w * w.w - z.x * w.x - z.y * w.y - z.z * w.z;
    r.x = z.w * w.x + z.x * w.w + z.y * w.z - z.z * w.y;
    r.y = z.w * w.y - z.x * w.z + z.y * w.w + z.z * w.x;
    r.z = z.w * w.z + z.x * w.y - z.y * w.x + z.z * w.w;
    return r;
}

// base.hpp:70-77 SE3Product: (qA qB, R(qA) tB + tA)
template <class T> inline Rigid<T> compose(const Rigid<T>& a, const Rigid<T>& b) {
    Rigid<T> r;
    r.q = mul(a.q, b.q);
    r.t = a.t + rotate(a.q, b.t);
    return r;
}

// Euler triple used all over the lidar / prior factors.  Array order is
// [yaw(Z), pitch(Y), roll(X)]  ("the real order of rpy is y p r", lidar_error.hpp:53).
template <class T> struct Ypr { T yaw, pitch, roll; };

// base.hpp:94-108 QuaternionToRPY / EigenQuaternionToRPY
template <class T> inline Ypr<T> to_ypr(const Quat<T>& e) {
    const T q0 = e.w, q1 = e.x, q2 = e.y, q3 = e.z;
    Ypr<T> r;
    r.yaw = atan2(T(2) * (q1 * q2 + q0 * q3), T(1) - T(2) * (q2 * q2 + q3 * q3));
    r.pitch = asin(T(2) * (q0 * q2 - q1 * q3));
    r.roll = atan2(T(2) * (q2 * q3 + q0 * q1), T(1) - T(2) * (q1 * q1 + q2 * q2));
    return r;
}

// base.hpp:110-132 RPYToQuaternion / RPYToEigenQuaternion  (R = Rz Ry Rx)
template <class T> inline Quat<T> from_ypr(const Ypr<T>& a) {
    const T hz = a.yaw / T(2), hy = a.pitch / T(2), hx = a.roll / T(2);
    const T cz = cos(hz), sz = sin(hz);
    const T cy = cos(hy), sy = sin(hy);
    const T cx = cos(hx), sx = sin(hx);
    Quat<T> q;
    q.w = cz * cy * cx + sz * sy * sx;
    q.x = cz * cy * sx - sz * sy * cx;
    q.y = cz * sy * cx + sz * cy * sx;
    q.z = sz * cy * cx - cz * sy * sx;
    return q;
}

// base.hpp:134-141 SE3ToRpyxyz : out[6] = [yaw pitch roll x y z]
template <class T> inline void to_rpyxyz(const Rigid<T>& a, T* out) {
    const Ypr<T> e = to_ypr(a.q);
    out[0] = e.yaw; out[1] = e.pitch; out[2] = e.roll;
    out[3] = a.t.x; out[4] = a.t.y; out[5] = a.t.z;
}

// base.hpp:143-150 RpyxyzToSE3
template <class T> inline Rigid<T> from_rpyxyz(const T* in) {
    Rigid<T> r;
    r.q = from_ypr(Ypr<T>{in[0], in[1], in[2]});
    r.t = Vec3<T>(in[3], in[4], in[5]);
    return r;
}

// ---------------------------------------------------------------------------------------
// Eigen::Quaterniond operations used by the IMU factor (double only).  Eigen is not
// vendored either; these are the textbook definitions Eigen 3.3 implements [upstream].
// ---------------------------------------------------------------------------------------
typedef Quat<double> Qd;
typedef Vec3<double> V3d;

// Eigen QuaternionBase::inverse(): conj / squaredNorm
inline Qd eig_inverse(const Qd& q) {
    const double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    return Qd(-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2);
}
// Eigen QuaternionBase::_transformVector: v + w*(2 qv x v) + qv x (2 qv x v)
inline V3d eig_rotate(const Qd& q, const V3d& v) {
    const V3d qv(q.x, q.y, q.z);
    V3d uv = cross(qv, v);
    uv = uv + uv;
    return v + uv * q.w + cross(qv, uv);
}
inline Qd eig_mul(const Qd& a, const Qd& b) { return mul(a, b); }  // Hamilton product

struct Mat3 {
    double m[3][3];
    Mat3() { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = 0.0; }
    static Mat3 identity() { Mat3 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.0; return r; }
};
inline Mat3 operator*(const Mat3& a, const Mat3& b) {
    Mat3 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0.0; for (int k = 0; k < 3; ++k) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; }
    return r;
}
inline Mat3 operator*(const Mat3& a, double s) { Mat3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] * s; return r; }
inline Mat3 operator+(const Mat3& a, const Mat3& b) { Mat3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
inline Mat3 operator-(const Mat3& a, const Mat3& b) { Mat3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j]; return r; }
inline Mat3 operator-(const Mat3& a) { return a * -1.0; }
inline V3d operator*(const Mat3& a, const V3d& v) {
    return V3d(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
               a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
               a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
inline Mat3 transpose(const Mat3& a) { Mat3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i]; return r; }

// Eigen QuaternionBase::toRotationMatrix()
inline Mat3 eig_matrix(const Qd& q) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    Mat3 r;
    r.m[0][0] = 1 - (tyy + tzz); r.m[0][1] = txy - twz;       r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz;       r.m[1][1] = 1 - (txx + tzz); r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy;       r.m[2][1] = tyz + twx;       r.m[2][2] = 1 - (txx + tyy);
    return r;
}

// /root/reference/src/lvio_fusion/include/lvio_fusion/utility.h:114-122 skew_symmetric
inline Mat3 skew(const V3d& v) {
    Mat3 r;
    r.m[0][1] = -v.z; r.m[0][2] = v.y;
    r.m[1][0] = v.z;  r.m[1][2] = -v.x;
    r.m[2][0] = -v.y; r.m[2][1] = v.x;
    return r;
}

}  // namespace oracle
