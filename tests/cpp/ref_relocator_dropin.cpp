// tests/cpp/ref_relocator_dropin.cpp -- TEST INFRASTRUCTURE ONLY: drop-in proof for the off-path solves, CPU
// only, development container only.
//
// The REFERENCE's src/relocator.cpp, compiled where it lies with its own factor header, runs Relocator::UpdateNewSubmap
// (src/relocator.cpp:247-282): one 4-parameter quaternion block under ceres::EigenQuaternionParameterization, one
// RelocateRError (7 residuals, pose_error.hpp:192-224) per keyframe of the new submap, DENSE_QR -- solved by the product shim's
// host LM -- then the rigid update of the submap.  The odometry of the submap is the relocated geometry seen through one
// rotation about the best frame; the solve has to find it and every keyframe has to land on its relocated pose.
#include "lvio_fusion/common.h"
#include <cstdio>
#define private public
#include "lvio_fusion/loop/relocator.h"
#undef private
#include "lvio_fusion/ceres/pose_error.hpp"
#include "lvio_fusion/manager.h"
#include "lvio_fusion/map.h"

const double epsilon = 1e-3;          // src/estimator.cpp:9-10
const int num_threads = 1;
namespace lvio_fusion { Matrix3d normalize_R(const Matrix3d&) { std::abort(); } }
using namespace lvio_fusion;

static Quaterniond rpy_q(double yaw, double pitch, double roll) {
    const Quaterniond qz(std::cos(0.5 * yaw), 0, 0, std::sin(0.5 * yaw)), qy(std::cos(0.5 * pitch), 0, std::sin(0.5 * pitch), 0), qx(std::cos(0.5 * roll), std::sin(0.5 * roll), 0, 0);
    return qz * qy * qx;
}

int main() {
    Camera::Create(718.856, 718.856, 607.1928, 185.2157, SE3d());          // Frame::Frame reads Camera::Get()->fx
    Camera::Create(718.856, 718.856, 607.1928, 185.2157, SE3d());
    const int n = 6, best = 2;
    std::vector<SE3d> truth(n);
    for (int k = 0; k < n; ++k) truth[k] = SE3d(rpy_q(0.8 + 0.05 * k, 0.01 * k, -0.005 * k), Vector3d(40 + 1.5 * k, -12 + 0.4 * k, 0.1 * k));
    const SE3d R0(rpy_q(0.05, 0.012, -0.007), Vector3d(0, 0, 0));          // the rotation the odometry is off by, about the best frame
    const SE3d base = SE3d(rpy_q(0.3, 0, 0), Vector3d(2.0, -1.0, 0.3)) * truth[best];        // where odometry believes the best frame is
    Frames submap; std::vector<Frame::Ptr> frames;
    for (int k = 0; k < n; ++k) {
        Frame::Ptr f = Frame::Create(); f->time = 300.0 + k;
        // odometry: the relocated geometry relative to the best frame, turned by R0^-1 about it (the best frame itself cannot turn
        // relative to itself: its own block pulls towards the identity, the other five towards R0)
        f->pose = k == best ? base : base * R0.inverse() * (truth[best].inverse() * truth[k]);
        Frame::Ptr old = Frame::Create(); old->time = 10.0 + k;
        old->pose = SE3d(rpy_q(0.75 + 0.05 * k, 0, 0), Vector3d(39 + 1.5 * k, -11.5 + 0.4 * k, 0));
        f->loop_closure = loop::LoopClosure::Ptr(new loop::LoopClosure());
        f->loop_closure->frame_old = old;
        f->loop_closure->relative_o_c = old->pose.inverse() * truth[k];   // what Mapping::Relocate / the matcher measured
        submap[f->time] = f; frames.push_back(f);
    }
    // the objective of relocator.cpp:256-262, evaluated with the reference's own functor
    std::vector<RelocateRError> blocks;
    // (the best frame's own pose is overwritten with the relocated one before the loop, :254, so its block reads base^-1 * relocated)
    for (int k = 0; k < n; ++k) blocks.emplace_back(truth[best].inverse() * truth[k], base.inverse() * (k == best ? truth[best] : frames[k]->pose));
    auto cost = [&](const Quaterniond& q) {
        const double r[4] = {q.x(), q.y(), q.z(), q.w()};
        double c = 0, res[7];
        for (auto& b : blocks) { b(r, res); for (double v : res) c += 0.5 * v * v; }
        return c;
    };
    alignas(Relocator) static unsigned char storage[sizeof(Relocator)];     // UpdateNewSubmap touches no member; the constructor would start the detector thread
    Relocator* relocator = reinterpret_cast<Relocator*>(storage);
    relocator->UpdateNewSubmap(frames[best], submap);
    const Quaterniond r = truth[best].unit_quaternion().conjugate() * frames[best]->pose.unit_quaternion();      // best pose = relocated * r (:271)
    const double c_star = cost(r), c_identity = cost(Quaterniond()), c_r0 = cost(R0.unit_quaternion());
    unsigned st = 99; auto u = [&]() { st = st * 1664525u + 1013904223u; return (double)(st >> 8) / 16777216.0 - 0.5; };
    double best_other = 1e300;
    for (int i = 0; i < 2000; ++i) {
        const double s = i % 2 ? 1e-2 : 1e-4;
        Quaterniond dq(1, s * u(), s * u(), s * u()); dq.normalize();
        best_other = std::min(best_other, cost(r * dq));
    }
    const double away_from_identity = 2 * r.vec().norm(), to_r0 = 2 * (R0.unit_quaternion().conjugate() * r).vec().norm();
    double rigid = 0;       // every other keyframe moved by the same transform as the best one
    const SE3d transform = frames[best]->pose * base.inverse();
    for (int k = 0; k < n; ++k) if (k != best) {
        const SE3d expect = transform * (base * R0.inverse() * (truth[best].inverse() * truth[k]));
        rigid = std::max(rigid, (expect.inverse() * frames[k]->pose).translation().norm());
    }
    printf("relocate cost %.9e identity %.9e r0 %.9e best_perturbed %.9e away %.4e to_r0 %.4e rigid %.3e\n", c_star, c_identity, c_r0, best_other, away_from_identity, to_r0, rigid);
    // optimal to within what Ceres' function tolerance (1e-6 relative, checked on the candidate step) leaves on the table
    return (c_star <= best_other * (1 + 5e-6) && c_star < c_identity * (1 - 5e-4) && c_star <= c_r0 * (1 + 5e-6) && away_from_identity > 0.02 && to_r0 < 0.03 && rigid < 1e-9) ? 0 : 1;
}
