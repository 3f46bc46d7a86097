// tests/cpp/test_shim.cpp -- drives the C++ shim (include/lvio_b200/*.h) the way the reference's backend does:
// parameter blocks are raw pointers into "Frame"/"Landmark"-owned memory, residual blocks come from the factor
// factories, ceres::Solve updates the memory in place.  Reads a problem dumped by tests/test_gpu_shim.py, writes
// the optimised parameters back for comparison with the CPU oracle.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include "lvio_b200/association.h"
#include "lvio_b200/factors.h"
#include "lvio_b200/types.h"

using lvb::SE3d; using lvb::Vector2d; using lvb::Vector3d;

static std::vector<double> rd(FILE* f, size_t n) { std::vector<double> v(n); if (n && fread(v.data(), 8, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }
static std::vector<int32_t> ri(FILE* f, size_t n) { std::vector<int32_t> v(n); if (n && fread(v.data(), 4, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }

// mirrors of the reference's objects that own parameter memory (frame.h, visual/landmark.h, imu/preintegration.h)
struct Frame { SE3d pose; Vector3d Vw, ba, bg; };
struct Landmark { double inv_depth; };
struct Coeffs { double v[4]; const double* data() const { return v; } };
struct Quat { Coeffs c; const Coeffs& coeffs() const { return c; } };
struct Mat15 { std::vector<double> v; const double* data() const { return v.data(); } };
struct Preintegration { Vector3d delta_p, delta_v, linearized_ba, linearized_bg; Quat delta_q; double sum_dt; Mat15 jacobian, covariance; };

static int run_ba(const char* in, const char* out, bool register_rig = true) {
    FILE* f = fopen(in, "rb"); if (!f) { perror(in); return 2; }
    std::vector<int32_t> h = ri(f, 10);          // np nv nr n0..n5 max_iter
    const int np = h[0], nv = h[1], nr = h[2];
    std::vector<double> cam = rd(f, 22), P = rd(f, 7 * (size_t)np), V = rd(f, 3 * (size_t)nv), R = rd(f, nr);
    static const int cs[6] = {5, 6, 5, 469, 8, 9}, is[6] = {3, 1, 1, 8, 2, 1};
    std::vector<double> C[6]; std::vector<int32_t> I[6];
    for (int k = 0; k < 6; ++k) { C[k] = rd(f, (size_t)h[3 + k] * cs[k]); I[k] = ri(f, (size_t)h[3 + k] * is[k]); }
    fclose(f);
    if (register_rig) lvb::Runtime::get().set_cameras(cam.data(), cam.data() + 11);          // once per process (Estimator); the context itself is per thread
    std::vector<Frame> frames(np); std::vector<Landmark> lms(nr);
    std::vector<Vector3d> vecs(nv);
    for (int i = 0; i < np; ++i) for (int k = 0; k < 7; ++k) frames[i].pose.v[k] = P[7 * i + k];
    for (int i = 0; i < nv; ++i) for (int k = 0; k < 3; ++k) vecs[i].v[k] = V[3 * i + k];
    for (int i = 0; i < nr; ++i) lms[i].inv_depth = R[i];
    lvb::Camera::Ptr cam0(new lvb::Camera()), cam1(new lvb::Camera());

    ceres::Problem problem;
    ceres::LossFunction* loss = new ceres::HuberLoss(1.0);                                   // backend.cpp:98
    ceres::LocalParameterization* local = new ceres::ProductParameterization(new ceres::EigenQuaternionParameterization(), new ceres::IdentityParameterization(3));
    for (int i = 0; i < np; ++i) problem.AddParameterBlock(frames[i].pose.data(), 7, local);  // backend.cpp:111
    for (int f2 = 0; f2 < h[3]; ++f2) {                                                       // backend.cpp:132-140
        const double* c = &C[0][5 * (size_t)f2]; const int32_t* ix = &I[0][3 * (size_t)f2];
        problem.AddParameterBlock(&lms[ix[0]].inv_depth, 1);
        problem.AddResidualBlock(lvio_fusion::TwoFrameReprojectionError::Create(Vector2d(c[0], c[1]), Vector2d(c[2], c[3]), cam0, cam1, c[4]), loss,
                                 &lms[ix[0]].inv_depth, frames[ix[1]].pose.data(), frames[ix[2]].pose.data());
    }
    for (int f2 = 0; f2 < h[4]; ++f2) {                                                       // backend.cpp:126-131
        const double* c = &C[1][6 * (size_t)f2];
        problem.AddResidualBlock(lvio_fusion::PoseOnlyReprojectionError::Create(Vector2d(c[0], c[1]), Vector3d(c[2], c[3], c[4]), cam0, c[5]), loss, frames[I[1][f2]].pose.data());
    }
    for (int f2 = 0; f2 < h[5]; ++f2) {                                                       // backend.cpp:119-125
        const double* c = &C[2][5 * (size_t)f2];
        problem.AddParameterBlock(&lms[I[2][f2]].inv_depth, 1);
        problem.AddResidualBlock(lvio_fusion::TwoCameraReprojectionError::Create(Vector2d(c[0], c[1]), Vector2d(c[2], c[3]), cam0, cam1, c[4]), loss, &lms[I[2][f2]].inv_depth);
    }
    for (int f2 = 0; f2 < h[6]; ++f2) {                                                       // backend.cpp:143-162
        const double* c = &C[3][469 * (size_t)f2]; const int32_t* ix = &I[3][8 * (size_t)f2];
        std::shared_ptr<Preintegration> pre(new Preintegration());
        pre->delta_p = Vector3d(c[0], c[1], c[2]); for (int k = 0; k < 4; ++k) pre->delta_q.c.v[k] = c[3 + k];
        pre->delta_v = Vector3d(c[7], c[8], c[9]); pre->linearized_ba = Vector3d(c[10], c[11], c[12]); pre->linearized_bg = Vector3d(c[13], c[14], c[15]); pre->sum_dt = c[16];
        pre->jacobian.v.assign(c + 17, c + 242); pre->covariance.v.assign(c + 242, c + 467);
        for (int b = 1; b < 8; ++b) if (b != 4) problem.AddParameterBlock(vecs[ix[b]].data(), 3);
        problem.AddResidualBlock(lvio_fusion::ImuError::Create(pre, /*column major*/ false), NULL, frames[ix[0]].pose.data(), vecs[ix[1]].data(), vecs[ix[2]].data(), vecs[ix[3]].data(),
                                 frames[ix[4]].pose.data(), vecs[ix[5]].data(), vecs[ix[6]].data(), vecs[ix[7]].data());
    }
    for (int f2 = 0; f2 < h[8]; ++f2) {                                                       // backend.cpp:175
        const double* c = &C[5][9 * (size_t)f2]; SE3d p; for (int k = 0; k < 7; ++k) p.v[k] = c[k];
        problem.AddResidualBlock(lvio_fusion::PoseError::Create(p, c[7], c[8]), NULL, frames[I[5][f2]].pose.data());
    }
    ceres::Solver::Options options;
    options.linear_solver_type = ceres::SPARSE_SCHUR;                                         // backend.cpp:207
    options.max_num_iterations = h[9];
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    fprintf(stderr, "%s\n", summary.BriefReport().c_str());
    if (!summary.IsSolutionUsable()) return 3;
    FILE* o = fopen(out, "wb"); if (!o) { perror(out); return 2; }
    for (int i = 0; i < np; ++i) fwrite(frames[i].pose.data(), 8, 7, o);
    for (int i = 0; i < nv; ++i) fwrite(vecs[i].data(), 8, 3, o);
    for (int i = 0; i < nr; ++i) fwrite(&lms[i].inv_depth, 8, 1, o);
    const double s[4] = {summary.initial_cost, summary.final_cost, (double)summary.num_successful_steps, (double)summary.num_residual_blocks};
    fwrite(s, 8, 4, o); fclose(o);
    return 0;
}

// ---- Mapping::Optimize shape (mapping.cpp:139-191) with the reference's object layout
struct PointXYZI { float x, y, z, pad0, intensity, pad1, pad2, pad3; };     // pcl::PointXYZI is 32 bytes
struct Cloud { std::vector<PointXYZI> pts; size_t size() const { return pts.size(); } const PointXYZI* data() const { return pts.data(); } bool empty() const { return pts.empty(); } };
struct LidarFeature { Cloud points_ground, points_surf; };
struct Weights { double visual = 71.8856, lidar_ground = 1, lidar_surf = 0.01; };
struct LFrame { SE3d pose; std::shared_ptr<LidarFeature> feature_lidar; Weights weights; std::vector<int> features_left; };

static int run_icp(const char* in, const char* out, bool resident = false) {
    FILE* f = fopen(in, "rb"); if (!f) { perror(in); return 2; }
    std::vector<int32_t> h = ri(f, 4);            // n_scan n_map mode n_features_left
    std::vector<double> fp = rd(f, 7), mp = rd(f, 7), e = rd(f, 6);
    std::vector<float> scan((size_t)h[0] * 4), map((size_t)h[1] * 4);
    if (fread(scan.data(), 4, scan.size(), f) != scan.size() || fread(map.data(), 4, map.size(), f) != map.size()) return 2;
    fclose(f);
    std::shared_ptr<LFrame> frame(new LFrame()), map_frame(new LFrame());
    for (int k = 0; k < 7; ++k) { frame->pose.v[k] = fp[k]; map_frame->pose.v[k] = mp[k]; }
    frame->feature_lidar.reset(new LidarFeature()); map_frame->feature_lidar.reset(new LidarFeature());
    Cloud& sc = h[2] == 0 ? frame->feature_lidar->points_ground : frame->feature_lidar->points_surf;
    Cloud& mc = h[2] == 0 ? map_frame->feature_lidar->points_ground : map_frame->feature_lidar->points_surf;
    for (int i = 0; i < h[0]; ++i) sc.pts.push_back(PointXYZI{scan[4 * i], scan[4 * i + 1], scan[4 * i + 2], 1.f, 0.f, 0.f, 0.f, 0.f});
    for (int i = 0; i < h[1]; ++i) mc.pts.push_back(PointXYZI{map[4 * i], map[4 * i + 1], map[4 * i + 2], 1.f, 0.f, 0.f, 0.f, 0.f});
    frame->features_left.resize(h[3]);
    lvio_fusion::FeatureAssociation association(0.2);
#ifndef LVB_NO_RESIDENT_MAP
    if (resident) {
        // the map frame comes from three keyframe clouds already on the device (Mapping::ToWorld put them there),
        // merged and hashed there (Mapping::BuildMapFrame); map_frame's own clouds are emptied to show they are no longer uploaded
        const double ident[7] = {0, 0, 0, 1, 0, 0, 0};
        const size_t third = (mc.pts.size() + 2) / 3;
        long long keys[3] = {7, 8, 9};
        for (int k = 0; k < 3; ++k) {
            Cloud part; part.pts.assign(mc.pts.begin() + std::min(mc.pts.size(), k * third), mc.pts.begin() + std::min(mc.pts.size(), (k + 1) * third));
            if (!association.AppendKeyframe(keys[k], part, part, ident)) { fprintf(stderr, "append failed: %s\n", lvb_last_error()); return 3; }
        }
        if (!association.BuildMapFrame(keys, 3, -1.0)) { fprintf(stderr, "build failed: %s\n", lvb_last_error()); return 3; }
        mc.pts.clear();
    }
#else
    (void)resident;
#endif
    double rpyxyz[6]; for (int k = 0; k < 6; ++k) rpyxyz[k] = e[k];
    ceres::Problem problem;
    const bool ok = h[2] == 0 ? association.ScanToMapWithGround(frame, map_frame, rpyxyz, problem) : association.ScanToMapWithSegmented(frame, map_frame, rpyxyz, problem);
    if (!ok) { fprintf(stderr, "association failed: %s\n", lvb_last_error()); return 3; }
    ceres::Solver::Options options;
    options.linear_solver_type = ceres::DENSE_QR; options.max_num_iterations = 4;              // mapping.cpp:160-161
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    fprintf(stderr, "%s\n", summary.BriefReport().c_str());
    if (!summary.IsSolutionUsable()) return 3;
    FILE* o = fopen(out, "wb"); if (!o) return 2;
    fwrite(rpyxyz, 8, 6, o);
    const double s[3] = {summary.initial_cost, summary.final_cost, (double)summary.num_residual_blocks};
    fwrite(s, 8, 3, o); fclose(o);
    return 0;
}

// Backend::BackendLoop, Backend::GlobalLoop and Relocator::DetectorLoop solve from their own threads (backend.cpp:19-20): N host
// threads build and solve the same window at once, each through its own per-thread context (lvb::Runtime), results to out.<k>
static int run_ba_threads(const char* in, const char* out, int n) {
    { FILE* f = fopen(in, "rb"); if (!f) { perror(in); return 2; } std::vector<int32_t> h = ri(f, 10); std::vector<double> cam = rd(f, 22); fclose(f);
      lvb::Runtime::get().set_cameras(cam.data(), cam.data() + 11); }
    std::vector<int> rc(n, -1); std::vector<const void*> rt(n, nullptr);
    std::vector<std::thread> th;
    for (int k = 0; k < n; ++k) th.emplace_back([&, k] { const std::string o = std::string(out) + "." + std::to_string(k); rc[k] = run_ba(in, o.c_str(), false); rt[k] = &lvb::Runtime::get(); });
    for (auto& t : th) t.join();
    for (int k = 0; k < n; ++k) { if (rc[k]) return rc[k]; for (int j = 0; j < k; ++j) if (rt[j] == rt[k]) { fprintf(stderr, "two threads shared a runtime\n"); return 4; } }
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 5 && std::string(argv[1]) == "ba_threads") return run_ba_threads(argv[2], argv[3], std::atoi(argv[4]));
    if (argc != 4) { fprintf(stderr, "usage: test_shim ba|icp|icp_resident in.bin out.bin | ba_threads in.bin out.bin n\n"); return 1; }
    return argv[1][0] == 'b' ? run_ba(argv[2], argv[3]) : run_icp(argv[2], argv[3], std::string(argv[1]) == "icp_resident");
}
