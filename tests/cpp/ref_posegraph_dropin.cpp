// tests/cpp/ref_posegraph_dropin.cpp -- TEST INFRASTRUCTURE ONLY: drop-in proof for the off-path solves, CPU
// only, development container only.
//
// The REFERENCE's src/pose_graph.cpp, compiled where it lies with its own factor header (include/lvio_fusion/ceres/
// pose_error.hpp: PoseGraphError and RError as AutoDiff functors), runs PoseGraph::BuildProblem + PoseGraph::Optimize
// (src/pose_graph.cpp:163-224) on a loop closure: <ceres/ceres.h> is the product shim, so the section pose graph -- two constant
// end poses, one 7-parameter block per section start with the quaternion parameterisation, PoseGraphError between neighbours,
// RError on every rotation -- is solved by include/lvio_b200/host_solver.h, and the reference's ForwardUpdate then carries the
// keyframes inside each section along.
#include "lvio_fusion/common.h"
#include <cstdio>
#include "lvio_fusion/loop/pose_graph.h"
#include "lvio_fusion/manager.h"
#include "lvio_fusion/map.h"

const double epsilon = 1e-3;          // src/estimator.cpp:9-10
const int num_threads = 1;
namespace lvio_fusion { Matrix3d normalize_R(const Matrix3d&) { std::abort(); } }

using namespace lvio_fusion;

static Quaterniond yaw_q(double y) { return Quaterniond(std::cos(0.5 * y), 0, 0, std::sin(0.5 * y)); }

int main() {
    Camera::Create(718.856, 718.856, 607.1928, 185.2157, SE3d());          // Frame::Frame reads Camera::Get()->fx
    Camera::Create(718.856, 718.856, 607.1928, 185.2157, SE3d());
    // a square-ish lap; odometry drifts in translation (linear in time) between the old keyframe and the loop start
    const int n = 41, i_old = 4, i_start = 36;
    std::vector<Frame::Ptr> frames; std::vector<SE3d> truth;
    for (int i = 0; i < n; ++i) {
        const double s = (double)i / (n - 1), ang = 2 * M_PI * s;
        Frame::Ptr f = Frame::Create();
        f->time = 100.0 + i;
        const SE3d T(yaw_q(ang + M_PI / 2), Vector3d(30 * std::cos(ang), 30 * std::sin(ang), 0.2 * std::sin(3 * ang)));
        truth.push_back(T);
        const double drift = i <= i_old ? 0.0 : std::min(1.0, (double)(i - i_old) / (i_start - i_old));
        f->pose = SE3d(T.unit_quaternion(), Vector3d(T.translation() + drift * Vector3d(1.5, -1.0, 0.4)));
        lvio_fusion::Map::Instance().InsertKeyFrame(f);
        frames.push_back(f);
    }
    Atlas sections;
    for (int i = i_old + 4; i < i_start; i += 6) { Section s; s.A = frames[i]->time; s.B = s.A; s.C = frames[std::min(i + 6, i_start)]->time; sections[s.A] = s; }
    Section submap; submap.A = frames[i_old]->time; submap.B = frames[i_start]->time; submap.C = frames[n - 1]->time;

    auto rmse = [&]() { double e = 0; int m = 0; for (int i = i_old + 1; i < i_start; ++i) { e += (frames[i]->pose.translation() - truth[i].translation()).squaredNorm(); ++m; } return std::sqrt(e / m); };
    const double before = rmse();
    const SE3d old_fixed = frames[i_old]->pose, start_fixed = truth[i_start];
    adapt::Problem problem;
    PoseGraph::Instance().BuildProblem(sections, submap, problem);          // relative poses are read off the drifted estimate here ...
    for (int i = i_start; i < n; ++i) frames[i]->pose = truth[i];            // ... then Relocator::UpdateNewSubmap moves the new submap (relocator.cpp:213-216)
    PoseGraph::Instance().Optimize(sections, submap, problem);
    const double after = rmse();
    double moved_const = 0;
    for (int k = 0; k < 7; ++k) moved_const += std::fabs(frames[i_old]->pose.data()[k] - old_fixed.data()[k]) + std::fabs(frames[i_start]->pose.data()[k] - start_fixed.data()[k]);
    double rot_err = 0;
    for (int i = i_old + 1; i < i_start; ++i) rot_err = std::max(rot_err, 2 * (truth[i].unit_quaternion().conjugate() * frames[i]->pose.unit_quaternion()).vec().norm());
    printf("sections %zu blocks %d residuals %d rmse_before %.6e rmse_after %.6e moved_const %.3e rot_err %.3e\n", sections.size(), problem.NumParameterBlocks(),
           problem.NumResidualBlocks(), before, after, moved_const, rot_err);
    return (after < 0.5 * before && moved_const == 0.0 && rot_err < 0.05) ? 0 : 1;
}
