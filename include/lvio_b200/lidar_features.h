// include/lvio_b200/lidar_features.h -- lidar feature extraction and IMU preintegration with the call shapes of
//   lvio_fusion::FeatureAssociation::Process  (/root/reference/src/lvio_fusion/src/association.cpp:88-94, ctor
//     include/lvio_fusion/lidar/association.h:21-26) and
//   lvio_fusion::imu::Preintegration::Append / Repropagate (include/lvio_fusion/imu/preintegration.h:27-40,
//     src/preintegration.cpp:129-142)
// over the C ABI (lvb_lidar_extract_features, lvb_imu_preintegrate).  Point types are duck-typed: any contiguous cloud
// (.size(), .data()) whose records start with float x, y, z (pcl::PointXYZ 16 B, pcl::PointXYZI 32 B) goes in; the outputs
// are written as (x, y, z, intensity) into the caller's point type through a small adapter.
#pragma once
#include <cstring>
#include <vector>
#include "ceres_shim.h"

namespace lvb {

struct PointXYZI16 { float x, y, z, intensity; };     // the packed record the C ABI returns

class LidarFeatureExtractor {
public:
    // same argument order as FeatureAssociation's constructor (association.h:21); `extrinsic` = Lidar::Get()->extrinsic.data()
    LidarFeatureExtractor(int num_scans, int horizon_scan, double ang_res_y, double ang_bottom, int ground_rows, double cycle_time,
                          double min_range, double max_range, double resolution, const double* extrinsic7 = nullptr) {
        lvb_lidar_default_config(&cfg_);
        cfg_.num_scans = num_scans; cfg_.horizon_scan = horizon_scan; cfg_.ang_res_y = ang_res_y; cfg_.ang_bottom = ang_bottom;
        cfg_.ground_rows = ground_rows; cfg_.cycle_time = cycle_time; cfg_.min_range = min_range; cfg_.max_range = max_range; cfg_.resolution = resolution;
        if (extrinsic7) std::memcpy(cfg_.extrinsic, extrinsic7, sizeof(cfg_.extrinsic));
        const size_t cap = (size_t)num_scans * horizon_scan;
        ground_.resize(cap); surf_.resize(cap);
    }
    const lvb_lidar_config& config() const { return cfg_; }

    // FeatureAssociation::Process: raw scan -> feature->points_ground / points_surf (robot frame).  Returns false and
    // keeps lvb_last_error() on failure (no CPU fallback).
    template <class CloudIn, class CloudOut>
    bool Process(const CloudIn& points, CloudOut& points_ground, CloudOut& points_surf) {
        Runtime& rt = Runtime::get();
        if (!rt.ensure()) return false;
        int32_t ng = 0, ns = 0;
        const int stride = (int)sizeof(points.data()[0]);
        if (lvb_lidar_extract_features(rt.ctx, &cfg_, points.data(), (int)points.size(), stride, &ground_[0].x, &ng, &surf_[0].x, &ns) != LVB_OK) {
            rt.error = lvb_last_error();
            return false;
        }
        unpack(ground_, ng, points_ground); unpack(surf_, ns, points_surf);
        return true;
    }

private:
    template <class CloudOut>
    static void unpack(const std::vector<PointXYZI16>& src, int n, CloudOut& dst) {
        dst.resize((size_t)n);
        for (int i = 0; i < n; ++i) { auto& q = dst[(size_t)i]; q.x = src[i].x; q.y = src[i].y; q.z = src[i].z; q.intensity = src[i].intensity; }
    }
    lvb_lidar_config cfg_;
    std::vector<PointXYZI16> ground_, surf_;
};

// One keyframe interval as Preintegration buffers it: dt_buf / acc_buf / gyr_buf, the seed measurement and the
// linearisation biases (preintegration.h:27-40).  preintegrate_batch(...) = Propagate over all intervals at once; calling it
// again with other biases is Repropagate (tools.cpp:87).  out[i] is the 469-double LVB_IMU record of interval i.
struct ImuInterval {
    std::vector<double> dt;               // n samples
    std::vector<double> acc, gyr;         // 3 n each
    double acc0[3], gyr0[3], ba[3], bg[3];
};
inline bool preintegrate_batch(const std::vector<ImuInterval>& intervals, const double noise4[4], std::vector<double>& out469) {
    Runtime& rt = Runtime::get();
    if (!rt.ensure()) return false;
    const int n = (int)intervals.size();
    std::vector<int32_t> first(n + 1, 0);
    for (int i = 0; i < n; ++i) first[i + 1] = first[i] + (int32_t)intervals[i].dt.size();
    std::vector<double> samples((size_t)first[n] * 7), a0(3 * (size_t)n), g0(3 * (size_t)n), ba(3 * (size_t)n), bg(3 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        const ImuInterval& v = intervals[i];
        for (size_t k = 0; k < v.dt.size(); ++k) {
            double* r = &samples[7 * ((size_t)first[i] + k)];
            r[0] = v.dt[k]; for (int c = 0; c < 3; ++c) { r[1 + c] = v.acc[3 * k + c]; r[4 + c] = v.gyr[3 * k + c]; }
        }
        for (int c = 0; c < 3; ++c) { a0[3 * i + c] = v.acc0[c]; g0[3 * i + c] = v.gyr0[c]; ba[3 * i + c] = v.ba[c]; bg[3 * i + c] = v.bg[c]; }
    }
    out469.assign((size_t)n * 469, 0.0);
    if (n == 0) return true;
    if (lvb_imu_preintegrate(rt.ctx, n, first.data(), samples.data(), a0.data(), g0.data(), ba.data(), bg.data(), noise4, out469.data()) != LVB_OK) {
        rt.error = lvb_last_error();
        return false;
    }
    return true;
}

}  // namespace lvb
