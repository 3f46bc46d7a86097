// oracle/icp.h -- scan-to-map association + point-to-plane ICP of the CPU oracle (TEST INFRASTRUCTURE ONLY).
// The association (gate, factor creation, loss, prior weight) and the plane factors are pinned against the reference's own
// association.cpp / lidar_error.hpp compiled in place (oracle/ref_assoc_harness.cpp, tests/golden/ref_assoc.npz); the exact 3-NN is
// pinned against FLANN's KDTreeSingleIndex as bundled with OpenCV (cv2.flann_Index, tests/test_flann_pin.py: same indices, bit-identical
// float32 distances); the order of exact distance ties and the LM schedule remain "parity unpinned".
//
//   association : /root/reference/src/lvio_fusion/src/association.cpp:270-384
//   driver      : /root/reference/src/lvio_fusion/src/mapping.cpp:139-191
//   factors     : lidar_error.hpp:42-110, pose_error.hpp:135-190 (see factors.h)
//
// kNN: pcl::KdTreeFLANN (FLANN KDTreeSingleIndex, exact, float32 L2_Simple) is NOT vendored.
// The oracle defines the result as: exact 3 nearest neighbours under d2 = dx*dx + dy*dy + dz*dz
// accumulated in float32 in that order WITHOUT fused multiply-add (compile with
// -ffp-contract=off), ordered ascending by (d2, index).  FLANN's true tie order is traversal
// dependent and unpinned; on tie-free inputs both definitions agree -- checked against the FLANN copy that ships inside OpenCV's
// Python module (PCL's own FLANN is not in the image; same index class, same search parameters).
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <limits>
#include <thread>
#include <vector>
#include "factors.h"
#include "lm.h"

namespace oracle {

void icp_parallel(int T, const std::function<void(int)>& fn);   // defined in oracle_capi.cpp (shares the BA worker pool)

struct Pt { float x, y, z; };

inline float dist2_f32(const Pt& a, const Pt& b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    float s = dx * dx;
    s = s + dy * dy;
    s = s + dz * dz;
    return s;
}

// association.cpp:287-294: tf = frame->pose.cast<float>() ; ceres::SE3TransformPoint<float>
inline Pt transform_f32(const double* pose, const Pt& p) {
    float tf[7];
    for (int i = 0; i < 7; ++i) tf[i] = (float)pose[i];
    const Rigid<float> T = load_rigid<float>(tf);
    const Vec3<float> q = apply(T, Vec3<float>(p.x, p.y, p.z));
    return Pt{q.x, q.y, q.z};
}

struct Knn3 { int32_t idx[3]; float d2[3]; };

inline void knn_insert(Knn3& k, int32_t i, float d) {
    // ascending by (d2, idx)
    if (d > k.d2[2] || (d == k.d2[2] && i > k.idx[2])) return;
    int pos = 2;
    while (pos > 0 && (d < k.d2[pos - 1] || (d == k.d2[pos - 1] && i < k.idx[pos - 1]))) { k.d2[pos] = k.d2[pos - 1]; k.idx[pos] = k.idx[pos - 1]; --pos; }
    k.d2[pos] = d; k.idx[pos] = i;
}
inline Knn3 knn_empty() { Knn3 k; for (int j = 0; j < 3; ++j) { k.idx[j] = std::numeric_limits<int32_t>::max(); k.d2[j] = std::numeric_limits<float>::infinity(); } return k; }

inline Knn3 knn3_brute(const Pt* map, int P, const Pt& q) {
    Knn3 k = knn_empty();
    for (int i = 0; i < P; ++i) knn_insert(k, i, dist2_f32(map[i], q));
    return k;
}

// Exact kd-tree (median split on the widest axis, leaf <= 16) -- used for the CPU baseline at
// sizes where brute force is hopeless.  Pruning is conservative in float32, so the result is the
// same set brute force returns.
struct KdTree {
    struct Node { int lo, hi, axis, left, right; float split; float bmin[3], bmax[3]; };
    std::vector<Node> nodes;
    std::vector<int32_t> order;
    const Pt* pts = nullptr;

    static float coord(const Pt& p, int a) { return a == 0 ? p.x : (a == 1 ? p.y : p.z); }
    void build(const Pt* p, int n) {
        pts = p; order.resize(n); for (int i = 0; i < n; ++i) order[i] = i;
        nodes.clear(); nodes.reserve(n / 4 + 16);
        if (n > 0) build_rec(0, n);
    }
    int build_rec(int lo, int hi) {
        Node nd; nd.lo = lo; nd.hi = hi; nd.left = nd.right = -1; nd.axis = 0; nd.split = 0;
        for (int a = 0; a < 3; ++a) { nd.bmin[a] = std::numeric_limits<float>::infinity(); nd.bmax[a] = -nd.bmin[a]; }
        for (int i = lo; i < hi; ++i) for (int a = 0; a < 3; ++a) { const float c = coord(pts[order[i]], a); nd.bmin[a] = std::min(nd.bmin[a], c); nd.bmax[a] = std::max(nd.bmax[a], c); }
        const int id = (int)nodes.size(); nodes.push_back(nd);
        if (hi - lo > 16) {
            int ax = 0; float ext = -1;
            for (int a = 0; a < 3; ++a) if (nd.bmax[a] - nd.bmin[a] > ext) { ext = nd.bmax[a] - nd.bmin[a]; ax = a; }
            const int mid = (lo + hi) / 2;
            std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi,
                             [&](int32_t a, int32_t b) { return coord(pts[a], ax) < coord(pts[b], ax); });
            nodes[id].axis = ax;
            const int l = build_rec(lo, mid), r = build_rec(mid, hi);
            nodes[id].left = l; nodes[id].right = r;
        }
        return id;
    }
    static float box_d2(const Node& n, const Pt& q) {
        float s = 0;
        for (int a = 0; a < 3; ++a) { const float c = coord(q, a); float d = 0; if (c < n.bmin[a]) d = n.bmin[a] - c; else if (c > n.bmax[a]) d = c - n.bmax[a]; s += d * d; }
        return s * 0.999f;  // conservative: never prune a box that could hold an equal-distance point
    }
    void search(int id, const Pt& q, Knn3& k) const {
        const Node& n = nodes[id];
        if (box_d2(n, q) > k.d2[2]) return;
        if (n.left < 0) { for (int i = n.lo; i < n.hi; ++i) knn_insert(k, order[i], dist2_f32(pts[order[i]], q)); return; }
        const float c = coord(q, n.axis);
        const Node& L = nodes[n.left];
        if (c <= L.bmax[n.axis]) { search(n.left, q, k); search(n.right, q, k); } else { search(n.right, q, k); search(n.left, q, k); }
    }
    Knn3 query(const Pt& q) const { Knn3 k = knn_empty(); if (!nodes.empty()) search(0, q, k); return k; }
};

// association.cpp:296-300 gate
inline bool gate(const Knn3& k, int P, float thr) {
    for (int j = 0; j < 3; ++j) if (!(k.idx[j] < P && k.d2[j] < thr)) return false;
    return true;
}

// ICP problem of one ScanToMapWith{Ground,Segmented} + Solve (3 free scalars).
struct IcpProblem {
    int mode = 0;                     // 0 = RPZ (ground), 1 = YXY (surf)
    std::vector<double> consts;       // n x 10 : p pa n weight
    int n = 0;
    double Twc1[7];
    double rpyxyz[6];                 // live array (association.cpp:316 passes the pointer)
    double huber_a = 0.0;
    double prior_weight = -1.0;       // < 0: no prior (relocate=true)
    double prior_target[3];
    int num_threads = 1;
    double H[9], g[3];
    double cand[3];

    static void free_index(int mode, int* f) { if (mode == 0) { f[0] = 1; f[1] = 2; f[2] = 5; } else { f[0] = 0; f[1] = 3; f[2] = 4; } }
    void get_free(double* x) const { int f[3]; free_index(mode, f); for (int i = 0; i < 3; ++i) x[i] = rpyxyz[f[i]]; }
    void set_free(const double* x) { int f[3]; free_index(mode, f); for (int i = 0; i < 3; ++i) rpyxyz[f[i]] = x[i]; }
    int dim() const { return 3; }

    double accumulate(const double* x, bool with_jac, double* Hout, double* gout) const {
        const int T = std::max(1, num_threads);
        std::vector<double> part((size_t)T * 13, 0.0);
        auto work = [&](int t) {
            double* a = &part[(size_t)t * 13];
            const int lo = (int)((int64_t)n * t / T), hi = (int)((int64_t)n * (t + 1) / T);
            for (int i = lo; i < hi; ++i) {
                double r, J[3];
                lidar_plane_eval(&consts[(size_t)i * 10], mode, Twc1, rpyxyz, x, &r, with_jac ? J : nullptr);
                double rho, sr; huber(huber_a, r * r, &rho, &sr);
                a[12] += 0.5 * rho;
                if (with_jac) { r *= sr; for (int k = 0; k < 3; ++k) J[k] *= sr;
                    for (int p = 0; p < 3; ++p) { a[9 + p] += J[p] * r; for (int q = 0; q < 3; ++q) a[3 * p + q] += J[p] * J[q]; } }
            }
        };
        icp_parallel(T, work);
        double acc[13] = {0};
        for (int t = 0; t < T; ++t) for (int k = 0; k < 13; ++k) acc[k] += part[(size_t)t * 13 + k];
        if (prior_weight >= 0.0) {
            typedef Dual<3> D;
            D f[3] = {D::seed(x[0], 0), D::seed(x[1], 1), D::seed(x[2], 2)}, res[3];
            icp_prior<D>(mode, prior_target, prior_weight, f, res);
            for (int k = 0; k < 3; ++k) {
                acc[12] += 0.5 * res[k].v * res[k].v;
                for (int p = 0; p < 3; ++p) { acc[9 + p] += res[k].d[p] * res[k].v; for (int q = 0; q < 3; ++q) acc[3 * p + q] += res[k].d[p] * res[k].d[q]; }
            }
        }
        if (with_jac) { for (int k = 0; k < 9; ++k) Hout[k] = acc[k]; for (int k = 0; k < 3; ++k) gout[k] = acc[9 + k]; }
        return acc[12];
    }
    double linearize(std::vector<double>& gv, std::vector<double>& hdiag) {
        double x[3]; get_free(x);
        const double c = accumulate(x, true, H, g);
        gv.assign(g, g + 3); hdiag = {H[0], H[4], H[8]};
        return c;
    }
    bool solve(const std::vector<double>& lambda, std::vector<double>& delta) {
        std::vector<double> A(H, H + 9), b = {-g[0], -g[1], -g[2]};
        for (int i = 0; i < 3; ++i) A[4 * i] += lambda[i];
        if (!cholesky_solve(A, 3, b)) return false;
        delta = b; return true;
    }
    double candidate_cost(const std::vector<double>& d) { double x[3]; get_free(x); for (int i = 0; i < 3; ++i) cand[i] = x[i] + d[i]; return accumulate(cand, false, nullptr, nullptr); }
    void accept() { set_free(cand); }
    double x_norm() const { double x[3]; get_free(x); return std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); }
    double step_norm() const { double x[3]; get_free(x); double s = 0; for (int i = 0; i < 3; ++i) s += (x[i] - cand[i]) * (x[i] - cand[i]); return std::sqrt(s); }
    double gradient_max_norm(const std::vector<double>& gv) const { return std::fmax(std::fabs(gv[0]), std::fmax(std::fabs(gv[1]), std::fabs(gv[2]))); }
};

}  // namespace oracle
