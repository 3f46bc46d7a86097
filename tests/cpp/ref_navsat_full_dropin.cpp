// tests/cpp/ref_navsat_full_dropin.cpp -- TEST INFRASTRUCTURE ONLY: drop-in proof for the off-path solves, CPU
// only, development container only.
//
// The REFERENCE's src/navsat.cpp (with src/pose_graph.cpp, map.cpp, frame.cpp, ...), compiled where it lies with its own factor
// headers (ceres/navsat_error.hpp, ceres/pose_error.hpp: AutoDiff functors), runs unchanged on the product shim's host LM:
//   Navsat::AddPoint -> Navsat::Initialize   (navsat.cpp:10-35,104-133: yaw first with x, y constant, then all three)
//   Navsat::Optimize(section)                (:135-156) = OptimizeBC on B (roll from NavsatRError, then NavsatRXError with
//                                            constant / bounded scalar blocks and HuberLoss(0.1), :192-262), OptimizeAB (pose
//                                            graph of PoseGraphError + TError with the quaternion parameterisation, :264-305),
//                                            OptimizeBC on every keyframe of B..C with only x free.
#include "lvio_fusion/common.h"
#include <cstdio>
#include <cstring>
#include "lvio_fusion/ceres/base.hpp"
#include "lvio_fusion/loop/pose_graph.h"
#include "lvio_fusion/manager.h"
#include "lvio_fusion/map.h"
#include "lvio_fusion/utility.h"

const double epsilon = 1e-3;          // src/estimator.cpp:9-10
const int num_threads = 1;
namespace lvio_fusion {
Matrix3d normalize_R(const Matrix3d&) { std::abort(); }
SE3d get_pose_from_two_points(const Vector3d&, const Vector3d&) { std::abort(); }       // Navsat::EstimatePose: not reached
// src/utility.cpp:27-40 (the rest of that file needs OpenCV): the conversions navsat.cpp calls, on top of the reference's base.hpp
void se32rpyxyz(const SE3d T, double* e) { ceres::EigenQuaternionToRPY(T.data(), e); std::memcpy(e + 3, T.data() + 4, 3 * sizeof(double)); }
SE3d rpyxyz2se3(const double* e) { double q[4]; ceres::RPYToEigenQuaternion(e, q); return SE3d(Quaterniond(q[3], q[0], q[1], q[2]), Vector3d(e[3], e[4], e[5])); }
}  // namespace lvio_fusion

using namespace lvio_fusion;

static unsigned g_seed = 4242u;
static double urand() { g_seed = g_seed * 1664525u + 1013904223u; return (double)(g_seed >> 8) / 16777216.0; }
static double nrand() { double s = 0; for (int i = 0; i < 12; ++i) s += urand(); return s - 6.0; }
static Quaterniond yaw_q(double y) { return Quaterniond(std::cos(0.5 * y), 0, 0, std::sin(0.5 * y)); }

int main() {
    Camera::Create(718.856, 718.856, 607.1928, 185.2157, SE3d());          // Frame::Frame reads Camera::Get()->fx
    Camera::Create(718.856, 718.856, 607.1928, 185.2157, SE3d());
    Navsat::Create(1.0, true);
    Navsat::Ptr nav = Navsat::Get();

    // the drive in the navsat (east-north-up) frame: straight, a left bend, straight again; keyframes every 2 m
    const int n = 70, iA = 25, iB = 35, iC = 65;
    std::vector<SE3d> enu(n);
    double x = 0, y = 0, heading = 0;
    for (int i = 0; i < n; ++i) {
        if (i > iA && i <= iB) heading += 0.08;
        enu[i] = SE3d(yaw_q(heading), Vector3d(x, y, 0.02 * i));
        x += 2.0 * std::cos(heading); y += 2.0 * std::sin(heading);
    }
    // the VIO world is the navsat frame seen through an unknown yaw and offset: what Navsat::Initialize estimates
    const double yaw_e = 0.3, x_e = 5.0, y_e = -3.0;
    const SE3d E(yaw_q(yaw_e), Vector3d(x_e, y_e, 0));
    std::vector<Frame::Ptr> frames; std::vector<SE3d> truth(n);
    for (int i = 0; i < n; ++i) {
        truth[i] = E * enu[i];
        Frame::Ptr f = Frame::Create();
        f->time = 50.0 + i;
        // odometry: exact up to B, where the bend left a heading error and a small pitch error (what the section correction
        // removes: a rigid motion of everything after B about B)
        const Quaterniond dq = i > iB ? yaw_q(0.03) * Quaterniond(std::cos(0.002), 0, std::sin(0.002), 0) : Quaterniond();
        const SE3d drift(dq, Vector3d(0, 0, 0));
        const SE3d rel = truth[iB].inverse() * truth[i];
        f->pose = i > iB ? truth[iB] * drift * rel : truth[i];
        f->last_keyframe = frames.empty() ? nullptr : frames.back();
        lvio_fusion::Map::Instance().InsertKeyFrame(f);
        frames.push_back(f);
        // a fix arrives half a second later (raw = position in the navsat frame + noise)
        const Vector3d p = enu[i].translation() + 0.5 * (enu[std::min(i + 1, n - 1)].translation() - enu[i].translation());
        nav->AddPoint(f->time + 0.5, p.x() + 0.05 * nrand(), p.y() + 0.05 * nrand(), p.z() + 0.05 * nrand(), Vector3d(0.01, 0.01, 1.0));
    }
    double e6[6]; se32rpyxyz(nav->extrinsic, e6);
    int with_fix = 0; for (auto& f : frames) with_fix += f->feature_navsat ? 1 : 0;
    printf("initialize done %d yaw %.6f x %.6f y %.6f fixes %d\n", (int)nav->initialized, e6[0], e6[3], e6[4], with_fix);
    // Initialize ran early (first 10 m, navsat.cpp:31-34): refine nothing here, but take the true extrinsic for the section test so
    // that the two checks stay independent
    nav->extrinsic = E;

    auto rms = [&](int lo, int hi) { double s = 0; for (int i = lo; i <= hi; ++i) s += (frames[i]->pose.translation() - truth[i].translation()).squaredNorm(); return std::sqrt(s / (hi - lo + 1)); };
    auto yaw_err = [&](int i) { return 2 * (truth[i].unit_quaternion().conjugate() * frames[i]->pose.unit_quaternion()).vec().norm(); };
    const double before = rms(iB + 1, iC), yaw_before = yaw_err(iC);
    Section s; s.A = frames[iA]->time; s.B = frames[iB]->time; s.C = frames[iC]->time; s.degree = 0.08 * (iB - iA) * 180 / M_PI;
    s.relative_B = frames[iB - 1]->pose.inverse() * frames[iB]->pose;
    nav->Optimize(s);
    printf("optimize rms_before %.6e rms_after %.6e yaw_before %.6e yaw_after %.6e ab_rms %.6e\n", before, rms(iB + 1, iC), yaw_before, yaw_err(iC), rms(iA, iB));
    return 0;
}
