// tests/cpp/compat_backend/shared/lvio_fusion/common.h -- TEST INFRASTRUCTURE ONLY (development container).
// common.h for compiling the reference's src/backend.cpp (and the small translation units it needs) where it lies against
// the PRODUCT shim: the third-party stand-ins of oracle/ref_compat (mini-Eigen, Sophus-shaped SE3d, cv / pcl names) plus the
// std headers and glog's LOG() that the real common.h brings in.
#pragma once
#include <bitset>
#include <condition_variable>
#include <functional>
#include <iostream>
#include <list>
#include <mutex>
#include <queue>
#include <thread>
#include "../../../../../oracle/ref_compat/lvio_fusion/common.h"
// the OpenCV names visual/extractor.h and visual/local_map.h declare members with (never used by the harness)
namespace cv {
struct Point2i { int x = 0, y = 0; };
struct DescriptorMatcher { static std::shared_ptr<DescriptorMatcher> create(const char*) { return std::make_shared<DescriptorMatcher>(); } };
template <class T> using Ptr = std::shared_ptr<T>;
}  // namespace cv
struct NullLog { template <class T> NullLog& operator<<(const T&) { return *this; } };
#define LOG(x) NullLog()
