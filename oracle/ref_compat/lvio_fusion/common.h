// oracle/ref_compat/lvio_fusion/common.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/lvio_fusion/common.h (which pulls Eigen, Sophus, OpenCV, PCL, glog -- none installed
// here): the mini-Eigen of ../mini_eigen.h, a Sophus-shaped SE3d and a few OpenCV / PCL names, just enough for the
// reference's sensor.h, visual/camera.h, imu/imu.h, imu/preintegration.h, utility.h, the factor headers under ceres/ and
// src/preintegration.cpp to compile unchanged.  The arithmetic of the functors is the reference's own code.
#pragma once
#include <cassert>
#include <chrono>
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../mini_eigen.h"
using namespace Eigen;

namespace Sophus {
struct SO3d { static constexpr int num_parameters = 4; };
// storage [qx qy qz qw tx ty tz], unit quaternion
struct SE3d {
    static constexpr int num_parameters = 7;
    double d[7];
    SE3d() : d{0, 0, 0, 1, 0, 0, 0} {}
    explicit SE3d(const double* p) { for (int i = 0; i < 7; ++i) d[i] = p[i]; }
    SE3d(const Quaterniond& q, const Vector3d& t) : d{q.x(), q.y(), q.z(), q.w(), t.x(), t.y(), t.z()} {}
    double* data() { return d; }
    const double* data() const { return d; }
    Quaterniond unit_quaternion() const { return Quaterniond(d[3], d[0], d[1], d[2]); }
    Vector3d translation() const { return Vector3d(d[4], d[5], d[6]); }
    Matrix3d rotationMatrix() const { return unit_quaternion().toRotationMatrix(); }
    SE3d inverse() const { const Quaterniond qi = unit_quaternion().conjugate(); return SE3d(qi, Vector3d(qi * Vector3d(-translation()))); }
    SE3d operator*(const SE3d& b) const { return SE3d(unit_quaternion() * b.unit_quaternion(), Vector3d(unit_quaternion() * b.translation() + translation())); }
    Vector3d operator*(const Vector3d& p) const { return Vector3d(unit_quaternion() * p + translation()); }
};
}  // namespace Sophus
typedef Sophus::SE3d SE3d;
typedef Sophus::SO3d SO3d;

typedef unsigned char uchar;
namespace cv {
struct Mat {};
template <class T> struct MatCommaInit { MatCommaInit& operator,(T) { return *this; } operator Mat() const { return Mat(); } };
template <class T> struct Mat_ : Mat { Mat_(int, int) {} };
template <class T> inline MatCommaInit<T> operator<<(const Mat_<T>&, T) { return MatCommaInit<T>(); }
struct Point2f { float x, y; Point2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} };
struct Point3f { float x, y, z; Point3f(float x_ = 0, float y_ = 0, float z_ = 0) : x(x_), y(y_), z(z_) {} };
struct KeyPoint { Point2f pt; float size; KeyPoint() : size(0) {} KeyPoint(Point2f p, float s) : pt(p), size(s) {} };
}  // namespace cv
namespace pcl { template <class PointT> struct PointCloud; }

class NotImplemented : public std::logic_error { public: NotImplemented() : std::logic_error("Function not yet implemented") {} };
