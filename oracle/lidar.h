// oracle/lidar.h -- lidar feature pipeline of the CPU oracle (TEST INFRASTRUCTURE ONLY, parity unpinned).
//
// Sequential restatement of the step *before* the scan-to-map path:
//   FeatureAssociation::Preprocess / Extract / AdjustDistortion / CalculateSmoothness / ExtractFeatures /
//   Sensor2Robot / SegmentGround         /root/reference/src/lvio_fusion/src/association.cpp:88-268
//   ImageProjection::Process and helpers /root/reference/src/lvio_fusion/src/projection.cpp:26-320
//   filter_points_by_distance            /root/reference/src/lvio_fusion/include/lvio_fusion/utility.h:70-96
//
// Pinned: lidar_preprocess + lidar_project_and_segment reproduce the reference's own filter_points_by_distance and
// ImageProjection::Process (src/projection.cpp compiled in place, oracle/ref_lidar_harness.cpp, tests/golden/ref_lidar.npz)
// bit for bit -- points, order, ranges, ground flags, columns, ring indices -- on the fixture and on four full 64 x 1800 sweeps.
// Toolchain-dependent detail: the unqualified abs / atan2 / sqrt calls of projection.cpp resolve to the float overloads when
// libstdc++'s <stdlib.h> / <math.h> wrappers are in scope (they are in the real build: Eigen -> <emmintrin.h> -> <mm_malloc.h>
// -> <stdlib.h>); with only <cmath>/<cstdlib> in scope abs(float) would be C's int abs and ~0.06 % of the cells change their
// ground flag.  The oracle (and the CUDA path) take abs as the float overload and evaluate atan2 / sqrt in *double*, rounded once
// to float: that is reproducible across CPU and GPU, gives the same bins as the reference's float evaluation on all of the
// above, and differs by at most one ulp in the sweep's start / end orientation.
// [upstream] behaviours that are not in the reference tree and cannot be verified in this container:
//   * pcl::VoxelGrid: keying ijk = floor(p * inverse_leaf) - min_b, idx = i + j*dx + k*dx*dy over the bounding box of the
//     input, output sorted by idx, centroid of all fields accumulated in float32.  PCL sorts with std::sort (order inside a
//     voxel unspecified); the oracle accumulates in ascending input index;
//   * pcl::RadiusOutlierRemoval: keep a point iff at least min_pts points (itself included) lie at squared float32
//     distance < radius^2 (FLANN radius search is strict);
//   * pcl::SACSegmentation (SACMODEL_PLANE, SAC_RANSAC, 100 iterations, probability 0.99, not random): mt19937 seeded
//     12345, rnd() = engine() >> 1 (boost::uniform_int<>(0, INT_MAX)), index sample by partial Fisher-Yates on a shuffled
//     index array that persists across iterations, adaptive iteration bound k; the output is the inlier set of the best
//     sampled plane (the least-squares refinement only touches the coefficients, which the reference discards).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <algorithm>
#include <limits>
#include <random>
#include <vector>
#include "icp.h"

namespace oracle {

struct LidarConfig {
    int num_scans, horizon_scan;
    double ang_res_y, ang_bottom;
    int ground_rows;
    double cycle_time, min_range, max_range, resolution;
    double extrinsic[7];
};

struct PointI { float x, y, z, intensity; };

struct Segmented {
    std::vector<PointI> pts;
    std::vector<float> range, curvature;
    std::vector<uint8_t> ground;
    std::vector<int32_t> col;
    std::vector<int32_t> start_ring, end_ring;
    float start_orientation = 0, end_orientation = 0, orientation_diff = 0;
};

// association.cpp:96-101 Preprocess: removeNaNFromPointCloud + filter_points_by_distance (utility.h:70-96)
inline std::vector<PointI> lidar_preprocess(const LidarConfig& c, const unsigned char* raw, int n, int stride) {
    std::vector<PointI> out;
    out.reserve(n);
    for (int i = 0; i < n; ++i) {
        const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        float d = p[0] * p[0];
        d = d + p[1] * p[1];
        d = d + p[2] * p[2];
        if ((double)d > c.min_range * c.min_range && (double)d < c.max_range * c.max_range) out.push_back(PointI{p[0], p[1], p[2], 0.0f});
    }
    return out;
}

// projection.cpp:26-320
inline Segmented lidar_project_and_segment(const LidarConfig& c, const std::vector<PointI>& points) {
    const int R = c.num_scans, W = c.horizon_scan;
    Segmented s;
    s.start_ring.assign(R, 0); s.end_ring.assign(R, 0);
    const float ang_res_x = (float)(360.0 / (double)(float)W);      // projection.h:38  360.0 / float(horizon_scan)
    const float ang_res_y = (float)c.ang_res_y, ang_bottom = (float)c.ang_bottom;
    const float alpha_x = (float)((double)ang_res_x / 180.0 * M_PI), alpha_y = (float)((double)ang_res_y / 180.0 * M_PI);
    const float theta = (float)(60.0 / 180.0 * M_PI);
    // FindStartEndAngle :42-55 (an empty cloud is undefined behaviour in the reference; defined here as orientation 0)
    if (!points.empty()) {
        s.start_orientation = (float)(-std::atan2((double)points.front().y, (double)points.front().x));
        s.end_orientation = (float)(-std::atan2((double)points.back().y, (double)points.back().x) + 2 * M_PI);
        if ((double)(s.end_orientation - s.start_orientation) > 3 * M_PI) s.end_orientation = (float)((double)s.end_orientation - 2 * M_PI);
        else if ((double)(s.end_orientation - s.start_orientation) < M_PI) s.end_orientation = (float)((double)s.end_orientation + 2 * M_PI);
        s.orientation_diff = s.end_orientation - s.start_orientation;
    }
    // ProjectPointCloud :57-101
    std::vector<float> range_mat((size_t)R * W, FLT_MAX);
    std::vector<PointI> full((size_t)R * W, PointI{std::numeric_limits<float>::quiet_NaN(), std::numeric_limits<float>::quiet_NaN(), std::numeric_limits<float>::quiet_NaN(), -1.0f});
    for (const PointI& p : points) {
        float xy2 = p.x * p.x; xy2 = xy2 + p.y * p.y;
        const float vertical_angle = (float)(std::atan2((double)p.z, std::sqrt((double)xy2)) * 180 / M_PI);
        const int row = (int)((vertical_angle + ang_bottom) / ang_res_y);
        if (row < 0 || row >= R) continue;
        const float horizon_angle = (float)(std::atan2((double)p.x, (double)p.y) * 180 / M_PI);
        int colm = (int)(-std::round(((double)horizon_angle - 90.0) / (double)ang_res_x) + (double)(W / 2));
        if (colm >= W) colm -= W;
        if (colm < 0 || colm >= W) continue;
        float r2 = xy2 + p.z * p.z;
        const float range = (float)std::sqrt((double)r2);
        range_mat[(size_t)row * W + colm] = range;
        PointI q = p;
        q.intensity = (float)((double)(float)row + (double)(float)colm / 10000.0);
        full[(size_t)row * W + colm] = q;
    }
    // RemoveGround :103-153
    std::vector<int8_t> ground_mat((size_t)R * W, 0);
    std::vector<int32_t> label((size_t)R * W, 0);
    for (int j = 0; j < W; ++j)
        for (int i = 0; i < c.ground_rows && i + 1 < R; ++i) {
            const size_t lo = (size_t)i * W + j, up = (size_t)(i + 1) * W + j;
            if (full[lo].intensity == -1 || full[up].intensity == -1) { ground_mat[lo] = -1; continue; }
            const float dx = full[up].x - full[lo].x, dy = full[up].y - full[lo].y, dz = full[up].z - full[lo].z;
            float h2 = dx * dx; h2 = h2 + dy * dy;
            const float angle = (float)(std::atan2((double)dz, std::sqrt((double)h2)) * 180 / M_PI);
            if (std::fabs(angle) <= 10) { ground_mat[lo] = 1; ground_mat[up] = 1; }
        }
    for (size_t e = 0; e < (size_t)R * W; ++e) if (ground_mat[e] == 1 || range_mat[e] == FLT_MAX) label[e] = -1;
    // Segment / LabelComponents :155-320 (breadth-first search, literal)
    const int OUTLIER = 999999;
    int label_count = 1;
    std::vector<int> qx((size_t)R * W), qy((size_t)R * W), ax((size_t)R * W), ay((size_t)R * W);
    const int nbr[4][2] = {{-1, 0}, {0, 1}, {0, -1}, {1, 0}};
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < W; ++j) {
            if (label[(size_t)i * W + j] != 0) continue;
            std::vector<bool> line_flag(R, false);
            int qs = 0, qe = 1, pushed = 1;
            qx[0] = i; qy[0] = j; ax[0] = i; ay[0] = j;
            while (qs < qe) {
                const int fx = qx[qs], fy = qy[qs];
                ++qs;
                label[(size_t)fx * W + fy] = label_count;
                for (int t = 0; t < 4; ++t) {
                    const int tx = fx + nbr[t][0];
                    int ty = fy + nbr[t][1];
                    if (tx < 0 || tx >= R) continue;
                    if (ty < 0) ty = W - 1;
                    if (ty >= W) ty = 0;
                    if (label[(size_t)tx * W + ty] != 0) continue;
                    const float ra = range_mat[(size_t)fx * W + fy], rb = range_mat[(size_t)tx * W + ty];
                    const float d1 = std::max(ra, rb), d2 = std::min(ra, rb);
                    const float alpha = nbr[t][0] == 0 ? alpha_x : alpha_y;
                    const float angle = (float)std::atan2((double)d2 * std::sin((double)alpha), (double)d1 - (double)d2 * std::cos((double)alpha));
                    if (angle > theta) {
                        qx[qe] = tx; qy[qe] = ty; ++qe;
                        label[(size_t)tx * W + ty] = label_count;
                        line_flag[tx] = true;
                        ax[pushed] = tx; ay[pushed] = ty; ++pushed;
                    }
                }
            }
            bool feasible = false;
            if (pushed >= 30) feasible = true;
            else if (pushed >= 5) {
                int cnt = 0;
                for (int r = 0; r < R; ++r) if (line_flag[r]) ++cnt;
                if (cnt >= 3) feasible = true;
            }
            if (feasible) ++label_count;
            else for (int k = 0; k < pushed; ++k) label[(size_t)ax[k] * W + ay[k]] = OUTLIER;
        }
    // extraction :162-203
    int num = 0;
    for (int i = 0; i < R; ++i) {
        s.start_ring[i] = num - 1 + 5;
        for (int j = 0; j < W; ++j) {
            const size_t e = (size_t)i * W + j;
            if (label[e] > 0 || ground_mat[e] == 1) {
                if (label[e] == OUTLIER) continue;
                s.ground.push_back(ground_mat[e] == 1 ? 1 : 0);
                s.col.push_back(j);
                s.range.push_back(range_mat[e]);
                s.pts.push_back(full[e]);
                ++num;
            }
        }
        s.end_ring[i] = num - 1 - 5;
    }
    return s;
}

// association.cpp:113-149 AdjustDistortion (relative time into the fractional part of intensity)
inline void lidar_adjust_distortion(const LidarConfig& c, Segmented& s) {
    bool half_passed = false;
    for (size_t i = 0; i < s.pts.size(); ++i) {
        PointI& p = s.pts[i];
        float ori = (float)(-std::atan2((double)p.y, (double)p.x));
        if (!half_passed) {
            if ((double)ori < (double)s.start_orientation - M_PI / 2) ori = (float)((double)ori + 2 * M_PI);
            else if ((double)ori > (double)s.start_orientation + M_PI * 3 / 2) ori = (float)((double)ori - 2 * M_PI);
            if ((double)(ori - s.start_orientation) > M_PI) half_passed = true;
        } else {
            ori = (float)((double)ori + 2 * M_PI);
            if ((double)ori < (double)s.end_orientation - M_PI * 3 / 2) ori = (float)((double)ori + 2 * M_PI);
            else if ((double)ori > (double)s.end_orientation + M_PI / 2) ori = (float)((double)ori - 2 * M_PI);
        }
        const float rel_time = (ori - s.start_orientation) / s.orientation_diff;
        p.intensity = (float)((double)(int)p.intensity + c.cycle_time * (double)rel_time);
    }
}

// association.cpp:151-166 CalculateSmoothness.  Entries outside [5, size-5) are never written by the reference
// (uninitialised `new float[]`); the oracle defines them as 0.
inline void lidar_smoothness(Segmented& s) {
    const int size = (int)s.pts.size();
    s.curvature.assign(size, 0.0f);
    const std::vector<float>& r = s.range;
    for (int i = 5; i < size - 5; ++i) {
        const float dr = (r[i + 5] - r[i - 5]) / 10;
        const float e[9] = {r[i + 4] - r[i - 5] - 9 * dr, r[i + 3] - r[i - 5] - 8 * dr, r[i + 2] - r[i - 5] - 7 * dr, r[i + 1] - r[i - 5] - 6 * dr,
                            r[i] - r[i - 5] - 5 * dr, r[i - 1] - r[i - 5] - 4 * dr, r[i - 2] - r[i - 5] - 3 * dr, r[i - 3] - r[i - 5] - 2 * dr,
                            r[i - 4] - r[i - 5] - 1 * dr};
        float acc = e[0] * e[0];
        for (int k = 1; k < 9; ++k) acc = acc + e[k] * e[k];
        const float cov = acc / 9;
        s.curvature[i] = cov * 10 / r[i];
    }
}

// association.cpp:186-213 feature split: ground flag first, else surf when curvature < 1
inline void lidar_split_features(const LidarConfig& c, const Segmented& s, std::vector<PointI>& ground, std::vector<PointI>& surf) {
    const float threshold = 1;
    for (int i = 0; i < c.num_scans; ++i)
        for (int j = 0; j < 6; ++j) {
            const int sp = (s.start_ring[i] * (6 - j) + s.end_ring[i] * j) / 6;
            const int ep = (s.start_ring[i] * (5 - j) + s.end_ring[i] * (j + 1)) / 6 - 1;
            if (sp >= ep) continue;
            for (int k = sp; k <= ep; ++k) {
                if (k < 0 || k >= (int)s.pts.size()) continue;      // the reference would read out of bounds here
                if (s.ground[k]) ground.push_back(s.pts[k]);
                else if (s.curvature[k] < threshold) surf.push_back(s.pts[k]);
            }
        }
}

// [upstream] pcl::VoxelGrid<PointXYZI>::applyFilter, leaf (lx = ly = lz), all fields averaged
inline std::vector<PointI> voxel_grid(const std::vector<PointI>& in, float leaf) {
    std::vector<PointI> out;
    if (in.empty()) return out;
    const float inv = 1.0f / leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const PointI& p : in) {
        if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
        mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
        mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    int min_b[3], div_b[3];
    for (int a = 0; a < 3; ++a) {
        min_b[a] = (int)std::floor(mn[a] * inv);
        const int max_b = (int)std::floor(mx[a] * inv);
        div_b[a] = max_b - min_b[a] + 1;
    }
    const long long cells = (long long)div_b[0] * div_b[1] * div_b[2];
    if (cells > (long long)std::numeric_limits<int32_t>::max()) return in;      // PCL: "Leaf size is too small", output = input
    std::vector<std::pair<int32_t, int32_t>> key;
    key.reserve(in.size());
    for (size_t i = 0; i < in.size(); ++i) {
        const PointI& p = in[i];
        if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
        const int i0 = (int)(std::floor(p.x * inv) - (float)min_b[0]);
        const int i1 = (int)(std::floor(p.y * inv) - (float)min_b[1]);
        const int i2 = (int)(std::floor(p.z * inv) - (float)min_b[2]);
        key.emplace_back(i0 + i1 * div_b[0] + i2 * div_b[0] * div_b[1], (int32_t)i);
    }
    std::sort(key.begin(), key.end());       // (idx, input index): the oracle's definition of the in-voxel order
    for (size_t a = 0; a < key.size();) {
        size_t b = a;
        float sx = 0, sy = 0, sz = 0, si = 0;
        while (b < key.size() && key[b].first == key[a].first) {
            const PointI& p = in[key[b].second];
            sx = sx + p.x; sy = sy + p.y; sz = sz + p.z; si = si + p.intensity;
            ++b;
        }
        const float cnt = (float)(b - a);
        out.push_back(PointI{sx / cnt, sy / cnt, sz / cnt, si / cnt});
        a = b;
    }
    return out;
}

// [upstream] pcl::RadiusOutlierRemoval (not negative): order preserved
inline std::vector<PointI> radius_outlier_removal(const std::vector<PointI>& in, double radius, int min_pts) {
    std::vector<PointI> out;
    const float r2 = (float)(radius * radius);
    for (size_t i = 0; i < in.size(); ++i) {
        int cnt = 0;
        const Pt a{in[i].x, in[i].y, in[i].z};
        for (size_t j = 0; j < in.size() && cnt < min_pts; ++j)
            if (dist2_f32(Pt{in[j].x, in[j].y, in[j].z}, a) < r2) ++cnt;
        if (cnt >= min_pts) out.push_back(in[i]);
    }
    return out;
}

// [upstream] pcl::SACSegmentation plane RANSAC as described in the header; returns the inliers in input order
struct PlaneHyp { float a, b, c, d; bool valid; };
inline bool plane_sample_good(const PointI& p0, const PointI& p1, const PointI& p2) {
    const float d1[3] = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z}, d2[3] = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z};
    const float q0 = d1[0] / d2[0], q1 = d1[1] / d2[1], q2 = d1[2] / d2[2];
    return (q0 != q1) || (q2 != q1);
}
inline PlaneHyp plane_from_sample(const PointI& p0, const PointI& p1, const PointI& p2) {
    PlaneHyp h{0, 0, 0, 0, false};
    if (!plane_sample_good(p0, p1, p2)) return h;
    const float u[3] = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z}, v[3] = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z};
    float n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    float nn = n[0] * n[0]; nn = nn + n[1] * n[1]; nn = nn + n[2] * n[2];
    const float len = std::sqrt(nn);
    n[0] = n[0] / len; n[1] = n[1] / len; n[2] = n[2] / len;
    float dd = n[0] * p0.x; dd = dd + n[1] * p0.y; dd = dd + n[2] * p0.z;
    h.a = n[0]; h.b = n[1]; h.c = n[2]; h.d = -1.0f * dd; h.valid = true;
    return h;
}
inline float plane_distance(const PlaneHyp& h, const PointI& p) {
    float s = h.a * p.x; s = s + h.b * p.y; s = s + h.c * p.z; s = s + h.d;
    return std::fabs(s);
}
inline std::vector<PointI> segment_ground(const std::vector<PointI>& in, double threshold) {
    const int N = (int)in.size();
    std::vector<PointI> out;
    if (N < 3) return out;
    const int max_iterations = 100;
    const double probability = 0.99;
    std::mt19937 eng(12345u);
    std::vector<int> shuffled(N);
    for (int i = 0; i < N; ++i) shuffled[i] = i;
    int iterations = 0, best_count = -std::numeric_limits<int>::max(), skipped = 0;
    const int max_skip = max_iterations * 10;
    double k = 1.0;
    PlaneHyp best{0, 0, 0, 0, false};
    const double log_probability = std::log(1.0 - probability);
    while ((double)iterations < k && skipped < max_skip) {
        int smp[3] = {-1, -1, -1};
        bool got = false;
        for (int tries = 0; tries < 1000 && !got; ++tries) {
            for (int i = 0; i < 3; ++i) std::swap(shuffled[i], shuffled[i + (int)((eng() >> 1) % (uint32_t)(N - i))]);
            smp[0] = shuffled[0]; smp[1] = shuffled[1]; smp[2] = shuffled[2];
            got = plane_sample_good(in[smp[0]], in[smp[1]], in[smp[2]]);
        }
        if (!got) break;
        const PlaneHyp h = plane_from_sample(in[smp[0]], in[smp[1]], in[smp[2]]);
        if (!h.valid) { ++skipped; continue; }
        int cnt = 0;
        for (const PointI& p : in) if ((double)plane_distance(h, p) < threshold) ++cnt;
        if (cnt > best_count) {
            best_count = cnt; best = h;
            const double w = (double)best_count / (double)N;
            double p_no_outliers = 1.0 - std::pow(w, 3.0);
            p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
            p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
            k = log_probability / std::log(p_no_outliers);
        }
        ++iterations;
        if (iterations > max_iterations) break;
    }
    if (!best.valid) return out;
    for (const PointI& p : in) if ((double)plane_distance(best, p) < threshold) out.push_back(p);
    return out;
}

// association.cpp:243-252 Sensor2Robot: float32 rigid transform by the lidar extrinsic
inline void sensor_to_robot(const LidarConfig& c, std::vector<PointI>& pts) {
    for (PointI& p : pts) { const Pt q = transform_f32(c.extrinsic, Pt{p.x, p.y, p.z}); p.x = q.x; p.y = q.y; p.z = q.z; }
}

// association.cpp:88-94,103-111 Process -> Extract; :186-241 ExtractFeatures
inline void lidar_extract_features(const LidarConfig& c, const unsigned char* raw, int n, int stride, std::vector<PointI>& ground, std::vector<PointI>& surf) {
    Segmented s = lidar_project_and_segment(c, lidar_preprocess(c, raw, n, stride));
    lidar_adjust_distortion(c, s);
    lidar_smoothness(s);
    ground.clear(); surf.clear();
    lidar_split_features(c, s, ground, surf);
    const float leaf = (float)(2 * c.resolution);
    surf = voxel_grid(surf, leaf);
    surf = radius_outlier_removal(surf, 4 * c.resolution, 4);
    ground = voxel_grid(ground, leaf);
    ground = segment_ground(ground, 0.1 * c.resolution);
    sensor_to_robot(c, ground);
    sensor_to_robot(c, surf);
}

}  // namespace oracle
