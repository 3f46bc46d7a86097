// oracle/ref_compat/lvio_fusion/common.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/lvio_fusion/common.h (which pulls Eigen, Sophus, OpenCV, PCL, glog -- none installed
// here): just enough of Eigen::Vector2d/Vector3d, Sophus::SE3d and cv::Mat for the reference's sensor.h, visual/camera.h and
// the factor headers to compile unchanged.  The group operations of SE3d (unit quaternion, Hamilton product) are only used
// by constructors (PoseGraphError: last_pose.inverse() * pose); the functors' arithmetic is the reference's own code.
#pragma once
#include <cassert>
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

struct Vector2d {
    double v[2];
    Vector2d() : v{0, 0} {}
    Vector2d(double x, double y) : v{x, y} {}
    double x() const { return v[0]; } double y() const { return v[1]; }
    double operator()(int i, int = 0) const { return v[i]; }
    const double* data() const { return v; }
};
struct Vector3d {
    double v[3];
    Vector3d() : v{0, 0, 0} {}
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; }
    double operator()(int i, int = 0) const { return v[i]; }
    double operator[](int i) const { return v[i]; }
    const double* data() const { return v; }
    Vector3d operator-(const Vector3d& o) const { return Vector3d(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3d operator+(const Vector3d& o) const { return Vector3d(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vector3d operator-() const { return Vector3d(-v[0], -v[1], -v[2]); }
    Vector3d cross(const Vector3d& o) const { return Vector3d(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]); }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    void normalize() { const double n = norm(); v[0] /= n; v[1] /= n; v[2] /= n; }      // Eigen: divides by the norm, no guard
};

namespace Sophus {
struct SO3d { static constexpr int num_parameters = 4; };
// storage [qx qy qz qw tx ty tz], unit quaternion
struct SE3d {
    static constexpr int num_parameters = 7;
    double d[7];
    SE3d() : d{0, 0, 0, 1, 0, 0, 0} {}
    explicit SE3d(const double* p) { for (int i = 0; i < 7; ++i) d[i] = p[i]; }
    double* data() { return d; }
    const double* data() const { return d; }
    Vector3d translation() const { return Vector3d(d[4], d[5], d[6]); }
    Vector3d rotate(const Vector3d& p) const {           // R(q) p for a unit quaternion
        const double x = d[0], y = d[1], z = d[2], w = d[3];
        double u0 = y * p.v[2] - z * p.v[1], u1 = z * p.v[0] - x * p.v[2], u2 = x * p.v[1] - y * p.v[0];
        u0 += u0; u1 += u1; u2 += u2;
        return Vector3d(p.v[0] + w * u0 + (y * u2 - z * u1), p.v[1] + w * u1 + (z * u0 - x * u2), p.v[2] + w * u2 + (x * u1 - y * u0));
    }
    SE3d inverse() const {
        SE3d r; r.d[0] = -d[0]; r.d[1] = -d[1]; r.d[2] = -d[2]; r.d[3] = d[3];
        const Vector3d t = r.rotate(-translation());
        r.d[4] = t.v[0]; r.d[5] = t.v[1]; r.d[6] = t.v[2];
        return r;
    }
    SE3d operator*(const SE3d& b) const {
        SE3d r;
        const double ax = d[0], ay = d[1], az = d[2], aw = d[3], bx = b.d[0], by = b.d[1], bz = b.d[2], bw = b.d[3];
        r.d[3] = aw * bw - ax * bx - ay * by - az * bz;
        r.d[0] = aw * bx + ax * bw + ay * bz - az * by;
        r.d[1] = aw * by - ax * bz + ay * bw + az * bx;
        r.d[2] = aw * bz + ax * by - ay * bx + az * bw;
        const Vector3d t = rotate(b.translation()) + translation();
        r.d[4] = t.v[0]; r.d[5] = t.v[1]; r.d[6] = t.v[2];
        return r;
    }
    Vector3d operator*(const Vector3d& p) const { return rotate(p) + translation(); }
};
}  // namespace Sophus
typedef Sophus::SE3d SE3d;
typedef Sophus::SO3d SO3d;

namespace cv {
struct Mat {};
template <class T> struct MatCommaInit { MatCommaInit& operator,(T) { return *this; } operator Mat() const { return Mat(); } };
template <class T> struct Mat_ : Mat { Mat_(int, int) {} };
template <class T> inline MatCommaInit<T> operator<<(const Mat_<T>&, T) { return MatCommaInit<T>(); }
}  // namespace cv

class NotImplemented : public std::logic_error { public: NotImplemented() : std::logic_error("Function not yet implemented") {} };
