// oracle/ref_compat/lvio_fusion/common.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/lvio_fusion/common.h (which pulls Eigen, Sophus, OpenCV, PCL, glog -- none installed
// here): the mini-Eigen of ../mini_eigen.h, a Sophus-shaped SE3d and a few OpenCV / PCL names, just enough for the
// reference's sensor.h, visual/camera.h, imu/imu.h, imu/preintegration.h, utility.h, the factor headers under ceres/ and
// src/preintegration.cpp to compile unchanged.  The arithmetic of the functors is the reference's own code.
#pragma once
// The real common.h pulls Eigen/Core, which on x86-64 includes <emmintrin.h> -> <xmmintrin.h> -> <mm_malloc.h> -> <stdlib.h>:
// in C++ that is libstdc++'s wrapper with `using std::abs;` (and OpenCV / PCL include <math.h>, whose wrapper exports the
// float overloads of the math functions).  The unqualified abs(float) of src/projection.cpp:129 therefore resolves to the
// floating-point overload in the real build; the same two C headers are included here so that it does in this one too.
#include <math.h>
#include <stdlib.h>
#include <cassert>
#include <cfloat>
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
struct SO3d {                                        // storage: the unit quaternion (x, y, z, w)
    static constexpr int num_parameters = 4;
    double d[4];
    SO3d() : d{0, 0, 0, 1} {}
    explicit SO3d(const Quaterniond& q) : d{q.x(), q.y(), q.z(), q.w()} {}
    double* data() { return d; }
    const double* data() const { return d; }
    Vector3d operator*(const Vector3d& p) const { return Vector3d(Quaterniond(d[3], d[0], d[1], d[2]) * p); }
    SO3d operator*(const SO3d& o) const { return SO3d(Quaterniond(d[3], d[0], d[1], d[2]) * Quaterniond(o.d[3], o.d[0], o.d[1], o.d[2])); }
    SO3d inverse() const { return SO3d(Quaterniond(d[3], -d[0], -d[1], -d[2])); }
};
struct SE3f { float d[7]; float* data() { return d; } const float* data() const { return d; } };
// storage [qx qy qz qw tx ty tz], unit quaternion
struct SE3d {
    static constexpr int num_parameters = 7;
    double d[7];
    SE3d() : d{0, 0, 0, 1, 0, 0, 0} {}
    explicit SE3d(const double* p) { for (int i = 0; i < 7; ++i) d[i] = p[i]; }
    SE3d(const Quaterniond& q, const Vector3d& t) : d{q.x(), q.y(), q.z(), q.w(), t.x(), t.y(), t.z()} {}
    SE3d(const SO3d& r, const Vector3d& t) : d{r.d[0], r.d[1], r.d[2], r.d[3], t.x(), t.y(), t.z()} {}
    double* data() { return d; }
    const double* data() const { return d; }
    Quaterniond unit_quaternion() const { return Quaterniond(d[3], d[0], d[1], d[2]); }
    SO3d so3() const { return SO3d(unit_quaternion()); }
    Vector3d translation() const { return Vector3d(d[4], d[5], d[6]); }
    Matrix3d rotationMatrix() const { return unit_quaternion().toRotationMatrix(); }
    SE3d inverse() const { const Quaterniond qi = unit_quaternion().conjugate(); return SE3d(qi, Vector3d(qi * Vector3d(-translation()))); }
    SE3d operator*(const SE3d& b) const { return SE3d(unit_quaternion() * b.unit_quaternion(), Vector3d(unit_quaternion() * b.translation() + translation())); }
    Vector3d operator*(const Vector3d& p) const { return Vector3d(unit_quaternion() * p + translation()); }
    template <class T> SE3f cast() const { SE3f r; for (int i = 0; i < 7; ++i) r.d[i] = (T)d[i]; return r; }      // frame->pose.cast<float>() (association.cpp:287)
};
}  // namespace Sophus
typedef Sophus::SE3d SE3d;
typedef Sophus::SO3d SO3d;

typedef unsigned char uchar;
// OpenCV names: camera.h keeps two cv::Mat members filled through the comma initialiser (never read here);
// src/projection.cpp uses cv::Mat(rows, cols, type, cv::Scalar::all(v)) and at<T>(i, j) for its range / label / ground images.
#define CV_8S 1
#define CV_32S 4
#define CV_32F 5
#define CV_32FC3 21
namespace cv {
struct Scalar { double v; static Scalar all(double x) { Scalar s; s.v = x; return s; } };
struct Mat {
    int rows = 0, cols = 0, type = 0;
    std::vector<unsigned char> buf;
    Mat() {}
    Mat(int r, int c, int t, const Scalar& s) : rows(r), cols(c), type(t), buf((size_t)r * c * elem(t)) {
        for (int i = 0; i < r * c; ++i) {
            if (t == CV_32F) reinterpret_cast<float*>(buf.data())[i] = (float)s.v;
            else if (t == CV_32S) reinterpret_cast<int*>(buf.data())[i] = (int)s.v;
            else reinterpret_cast<signed char*>(buf.data())[i] = (signed char)s.v;
        }
    }
    static int elem(int t) { return t == CV_8S ? 1 : (t == CV_32FC3 ? 12 : 4); }
    template <class T> T& at(int i, int j) { return reinterpret_cast<T*>(buf.data())[(size_t)i * cols + j]; }
    // Frame::GetObservation (src/frame.cpp:45-77, the reinforcement-learning observation; compiled, never run here)
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t, Scalar::all(0)); }
    template <class T> T* begin() { return reinterpret_cast<T*>(buf.data()); }
    template <class T> T* end() { return reinterpret_cast<T*>(buf.data() + buf.size()); }
    Mat reshape(int, int) const { return *this; }
    template <class T> operator std::vector<T>() const { const T* p = reinterpret_cast<const T*>(buf.data()); return std::vector<T>(p, p + buf.size() / sizeof(T)); }
};
struct Vec3f { float v[3]; float& operator[](int i) { return v[i]; } };
template <class T> struct MatCommaInit { MatCommaInit& operator,(T) { return *this; } operator Mat() const { return Mat(); } };
template <class T> struct Mat_ : Mat { Mat_(int, int) {} };
template <class T> inline MatCommaInit<T> operator<<(const Mat_<T>&, T) { return MatCommaInit<T>(); }
struct Point2f { float x, y; Point2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} };
struct Point3f { float x, y, z; Point3f(float x_ = 0, float y_ = 0, float z_ = 0) : x(x_), y(y_), z(z_) {} };
struct KeyPoint { Point2f pt; float size; KeyPoint() : size(0) {} KeyPoint(Point2f p, float s) : pt(p), size(s) {} };
}  // namespace cv
// PCL names: point records and a cloud that is a vector with a header
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0, pad = 0; };
struct PointXYZI {                     // PCL layout: data[4] aliases x y z, then intensity in its own 16-byte lane
    union { float data[4]; struct { float x, y, z; }; };
    union { struct { float intensity; }; float data_c[4]; };
    PointXYZI() { data[0] = data[1] = data[2] = 0; data[3] = 1; data_c[0] = data_c[1] = data_c[2] = data_c[3] = 0; }
};
struct PointXYZRGB {                   // PCL layout: x y z pad, then b g r a packed into one 32-bit word
    float x = 0, y = 0, z = 0, pad0 = 0;
    union { struct { unsigned char b, g, r, a; }; unsigned int rgba; };
    float pad1 = 0, pad2 = 0, pad3 = 0;
    PointXYZRGB() : rgba(0xff000000u) {}
};
struct PCLHeader { unsigned int seq = 0; unsigned long long stamp = 0; std::string frame_id; };
inline void copy_intensity(const PointXYZ&, PointXYZI&) {}
inline void copy_intensity(const PointXYZI& p, PointXYZI& q) { q.intensity = p.intensity; }
inline void copy_intensity(const PointXYZI&, PointXYZ&) {}
inline void copy_intensity(const PointXYZ&, PointXYZ&) {}
inline void copy_intensity(const PointXYZRGB& p, PointXYZRGB& q) { q.rgba = p.rgba; }
template <class PointT> struct PointCloud {
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    PCLHeader header;
    std::vector<PointT> points;
    unsigned int width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); width = (unsigned int)points.size(); height = 1; return *this; }
    void push_back(const PointT& p) { points.push_back(p); width = (unsigned int)points.size(); height = 1; }
    void clear() { points.clear(); width = height = 0; }
    PointT& operator[](size_t i) { return points[i]; }
    const PointT& operator[](size_t i) const { return points[i]; }
    typename std::vector<PointT>::iterator begin() { return points.begin(); }
    typename std::vector<PointT>::iterator end() { return points.end(); }
    typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
    typename std::vector<PointT>::const_iterator end() const { return points.end(); }
    template <class It> void insert(typename std::vector<PointT>::iterator pos, It a, It b) { points.insert(pos, a, b); width = (unsigned int)points.size(); height = 1; }
    PointCloud operator+(const PointCloud& o) const { PointCloud r = *this; r.points.insert(r.points.end(), o.points.begin(), o.points.end()); r.width = (unsigned int)r.points.size(); r.height = 1; return r; }
};
}  // namespace pcl
typedef pcl::PointXYZ Point3;
typedef pcl::PointCloud<Point3> Point3Cloud;
typedef pcl::PointXYZI PointI;
typedef pcl::PointCloud<PointI> PointICloud;
typedef pcl::PointXYZRGB PointRGB;
typedef pcl::PointCloud<PointRGB> PointRGBCloud;

namespace boost { template <class T, class... A> inline std::shared_ptr<T> make_shared(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); } }

extern const double epsilon;      // src/config.cpp in the reference
extern const int num_threads;

class NotImplemented : public std::logic_error { public: NotImplemented() : std::logic_error("Function not yet implemented") {} };
