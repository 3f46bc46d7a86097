// include/lvio_b200/types.h -- plain stand-ins for the Eigen / Sophus value types the reference passes to the
// factor factories, for builds (like this repository's tests) where Eigen and Sophus are not installed.
// With the real libraries present the factories in factors.h take Eigen::Vector2d / Sophus::SE3d directly.
#pragma once
#include <memory>

namespace lvb {
struct Vector2d { double v[2]; Vector2d(double x = 0, double y = 0) : v{x, y} {} const double* data() const { return v; } double* data() { return v; } };
struct Vector3d { double v[3]; Vector3d(double x = 0, double y = 0, double z = 0) : v{x, y, z} {} const double* data() const { return v; } double* data() { return v; } };
// Sophus::SE3d::data() layout: unit quaternion (x, y, z, w) then translation
struct SE3d { double v[7]; SE3d() : v{0, 0, 0, 1, 0, 0, 0} {} const double* data() const { return v; } double* data() { return v; } };
struct Camera { double fx, fy, cx, cy; SE3d extrinsic; typedef std::shared_ptr<Camera> Ptr; };
}  // namespace lvb
