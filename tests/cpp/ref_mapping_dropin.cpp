// tests/cpp/ref_mapping_dropin.cpp -- TEST INFRASTRUCTURE ONLY: drop-in proof for SURVEY 8(a7/a8), built in the development
// container (the reference tree is needed to compile it); the prebuilt binaries travel in oracle/_ref/.
//
// The REFERENCE's src/mapping.cpp (Mapping::Optimize, BuildMapFrame, MergeScan, ToWorld) is compiled where it lies and
// runs unchanged: it builds the map frame from the last three lidar keyframes, calls
// association_->ScanToMapWithGround / ScanToMapWithSegmented, adapt::Solve (DENSE_QR, 4 iterations) and rpyxyz2se3
// (src/mapping.cpp:139-194).  The two association members are the only thing replaced, exactly as INTEGRATION.md section 3
// prescribes: they forward to the product's lvio_b200/association.h (one ScanToMapCost per call, solved by the shim's
// ceres::Solve through lvb_icp_scan_to_map).  Everything else FeatureAssociation owns (SegmentGround, ...) is the
// reference's src/association.cpp, compiled in place with its own two ScanToMap members renamed out of the way
// (-DScanToMapWithGround=ScanToMapWithGround_ref ...; PCL = the oracle-backed adapters of oracle/ref_compat/pcl_standin.h).
//
// Builds (tests/test_zz_ref_backend_dropin.py): lvb_* bound to liblvio_b200.so (CUDA), or renamed to the oracle's orc_*; and
// -DDROPIN_REFERENCE_ASSOCIATION: nothing replaced -- the reference's own ScanToMapWith* (kd-tree adapter, one
// LidarPlaneError AutoDiff block per accepted point, PoseErrorRPZ / YXY prior) solved by the shim's host LM
// (include/lvio_b200/host_solver.h), the cross-check of the fused device solve.
#include "lvio_fusion/common.h"
#include <cstdio>
#include <cstring>
#include "lvio_fusion/lidar/lidar.h"
#include "lvio_fusion/lidar/mapping.h"
#include "lvio_fusion/loop/pose_graph.h"
#include "lvio_fusion/map.h"
#include "lvio_fusion/utility.h"
#include "lvio_fusion/visual/camera.h"
#ifndef DROPIN_REFERENCE_ASSOCIATION
#define LVB_ASSOCIATION_CLASS ScanToMapDevice
#include "lvio_b200/association.h"
#endif

const double epsilon = 1e-3;          // src/estimator.cpp:9-10
const int num_threads = 1;

namespace lvio_fusion {
#ifndef DROPIN_REFERENCE_ASSOCIATION
// the two members of the reference's FeatureAssociation forward to the device class
static ScanToMapDevice& device() { static ScanToMapDevice d(Lidar::Get()->resolution); return d; }
void FeatureAssociation::ScanToMapWithGround(Frame::Ptr frame, Frame::Ptr map_frame, double* para, adapt::Problem& problem, bool relocate) {
    device().ScanToMapWithGround(frame, map_frame, para, problem, relocate);
}
void FeatureAssociation::ScanToMapWithSegmented(Frame::Ptr frame, Frame::Ptr map_frame, double* para, adapt::Problem& problem, bool relocate) {
    device().ScanToMapWithSegmented(frame, map_frame, para, problem, relocate);
}
#endif
// named by the translation units above, defined in ones the harness does not link
static int g_forward_updates = 0;
void PoseGraph::ForwardUpdate(SE3d, double, bool) { ++g_forward_updates; }       // src/pose_graph.cpp: moves the later keyframes along
Matrix3d normalize_R(const Matrix3d&) { std::abort(); }
// src/utility.cpp:27-40 (the rest of that file needs OpenCV's optical flow and SVD): the two conversions mapping.cpp calls,
// on top of the reference's own base.hpp helpers they wrap
void se32rpyxyz(const SE3d T, double* e) { ceres::EigenQuaternionToRPY(T.data(), e); std::memcpy(e + 3, T.data() + 4, 3 * sizeof(double)); }
SE3d rpyxyz2se3(const double* e) { double q[4]; ceres::RPYToEigenQuaternion(e, q); return SE3d(Quaterniond(q[3], q[0], q[1], q[2]), Vector3d(e[3], e[4], e[5])); }
}  // namespace lvio_fusion

using namespace lvio_fusion;

static unsigned g_seed = 77u;
static double urand() { g_seed = g_seed * 1664525u + 1013904223u; return (double)(g_seed >> 8) / 16777216.0; }
static double nrand() { double s = 0; for (int i = 0; i < 12; ++i) s += urand(); return s - 6.0; }

static Quaterniond rpy_q(double yaw, double pitch, double roll) {
    const Quaterniond qz(std::cos(0.5 * yaw), 0, 0, std::sin(0.5 * yaw)), qy(std::cos(0.5 * pitch), 0, std::sin(0.5 * pitch), 0), qx(std::cos(0.5 * roll), std::sin(0.5 * roll), 0, 0);
    return qz * qy * qx;
}

// One keyframe's lidar features in the ROBOT frame: ground returns on the plane z = 0 of the world (slightly rolling road),
// "surf" returns on two walls and a few pillars.  Every keyframe samples its own points of the same surfaces.
static void make_features(const SE3d& pose, lidar::Feature::Ptr f) {
    const SE3d inv = pose.inverse();
    auto push = [&](PointICloud& cloud, const Vector3d& pw, float ring) {
        const Vector3d pb = inv * pw;
        if (pb.norm() < 3.0 || pb.norm() > 28.0) return;
        PointI p; p.x = (float)pb.x(); p.y = (float)pb.y(); p.z = (float)pb.z(); p.intensity = ring;
        cloud.push_back(p);
    };
    const Vector3d c = pose.translation();
    for (int i = 0; i < 1500; ++i) {
        const double x = c.x() + 50 * (urand() - 0.5), y = c.y() + 30 * (urand() - 0.5);
        push(f->points_ground, Vector3d(x, y, 0.02 * std::sin(0.2 * x) + 0.01 * nrand()), (float)(i % 16));
    }
    for (int i = 0; i < 1200; ++i) {
        const double x = c.x() + 50 * (urand() - 0.5), z = 0.3 + 3.0 * urand();
        const int wall = i % 3;
        if (wall == 0) push(f->points_surf, Vector3d(x, 9.0 + 0.01 * nrand(), z), (float)(20 + i % 30));
        else if (wall == 1) push(f->points_surf, Vector3d(x, -7.5 + 0.01 * nrand(), z), (float)(20 + i % 30));
        else { const int k = (int)(urand() * 6); push(f->points_surf, Vector3d(6.0 * k + 3.0 + 0.01 * nrand(), -3.0 + 6.0 * urand(), z), (float)(20 + i % 30)); }   // cross walls: constrain x
    }
}

int main(int argc, char** argv) {
    const char* dump = argc > 1 ? argv[1] : nullptr;
    Camera::Create(718.856, 718.856, 607.1928, 185.2157, SE3d());          // Frame::Frame reads Camera::Get()->fx (src/frame.cpp:13)
    Camera::Create(718.856, 718.856, 607.1928, 185.2157, SE3d());
    Lidar::Create(0.2, SE3d());

    Mapping mapping;
    mapping.SetFeatureAssociation(FeatureAssociation::Ptr(new FeatureAssociation(64, 1800, 0.427, 24.9, 60, 0.1, 5.0, 30.0, 0, 0.5)));
    std::vector<Frame::Ptr> frames;
    std::vector<SE3d> truth;
    for (int k = 0; k < 6; ++k) {
        Frame::Ptr f = Frame::Create();
        f->time = 20.0 + 0.5 * k;
        f->pose = SE3d(rpy_q(0.03 * k, 0.004 * k, -0.003 * k), Vector3d(1.2 * k, 0.1 * k, 1.7 + 0.01 * k));
        f->feature_lidar = lidar::Feature::Create();
        make_features(f->pose, f->feature_lidar);
        lvio_fusion::Map::Instance().InsertKeyFrame(f);
        frames.push_back(f); truth.push_back(f->pose);
    }
    for (int k = 0; k < 3; ++k) mapping.ToWorld(frames[k]);               // the map so far (Mapping::ToWorld(Frame::Ptr), :208-222)
    // keyframes 3..5 come from the window solve with some error; Mapping::Optimize registers each against the last three and
    // adds it to the map in turn (:139-194)
    Frames active;
    for (int k = 3; k < 6; ++k) {
        const Quaterniond dq = rpy_q(0.012 * nrand(), 0.006 * nrand(), 0.006 * nrand());
        frames[k]->pose = SE3d(truth[k].unit_quaternion() * dq, Vector3d(truth[k].translation() + Vector3d(0.12 * nrand(), 0.12 * nrand(), 0.05 * nrand())));
        active[frames[k]->time] = frames[k];
    }
    auto report = [&](const char* tag) {
        for (int k = 3; k < 6; ++k) {
            const SE3d d = truth[k].inverse() * frames[k]->pose;
            printf("%s kf %d err_t %.9e err_r %.9e\n", tag, k, d.translation().norm(), 2 * d.unit_quaternion().vec().norm());
        }
    };
    report("before");
    mapping.Optimize(active);
    report("after");
    printf("map clouds %zu forward_updates %d ground0 %zu surf0 %zu\n", mapping.pointclouds_surf.size(), g_forward_updates,
           frames[3]->feature_lidar->points_ground.size(), frames[3]->feature_lidar->points_surf.size());
    // Mapping::Relocate (:246-300): the loop-closure registration -- four rounds of both scan-to-map solves without the prior
    // (relocate = true), against the map frame around an OLD keyframe, scored from Summary::num_residual_blocks_reduced and
    // Summary::final_cost (the only place the reference reads a Summary)
    Frame::Ptr old_frame = frames[1], current = frames[4];
    current->loop_closure = loop::LoopClosure::Ptr(new loop::LoopClosure());
    current->loop_closure->frame_old = old_frame;
    const SE3d true_rel = truth[1].inverse() * truth[4];
    current->loop_closure->relative_o_c = SE3d(true_rel.unit_quaternion() * rpy_q(0.02, -0.004, 0.003), Vector3d(true_rel.translation() + Vector3d(0.15, -0.1, 0.04)));
    SE3d relative_o_c;
    const int score = mapping.Relocate(old_frame, current, relative_o_c);
    const SE3d d_rel = (old_frame->pose * true_rel).inverse() * (old_frame->pose * relative_o_c);
    printf("relocate score %d err_t %.9e err_r %.9e\n", score, d_rel.translation().norm(), 2 * d_rel.unit_quaternion().vec().norm());
    if (dump) {
        FILE* f = fopen(dump, "wb");
        for (int k = 3; k < 6; ++k) fwrite(frames[k]->pose.data(), sizeof(double), 7, f);
        fwrite(relative_o_c.data(), sizeof(double), 7, f);
        const double sc = score; fwrite(&sc, sizeof(double), 1, f);
        fclose(f);
    }
    return 0;
}
