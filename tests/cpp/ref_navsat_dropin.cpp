// tests/cpp/ref_navsat_dropin.cpp -- drop-in proof for the off-path solves, CPU only, development container only.
// The REFERENCE's own navsat cost functors (include/lvio_fusion/ceres/navsat_error.hpp + base.hpp, compiled where they lie,
// Eigen / Sophus replaced by the stand-ins of oracle/ref_compat) are created through their own Create() -- which names
// ceres::AutoDiffCostFunction -- and solved by ceres::Solve, both coming from the product shim (include/lvio_b200).  The
// problem is the one Navsat::Initialize builds (src/navsat.cpp:104-129): yaw first with x, y constant, then all three.
#include <cmath>
#include <cstdio>
#include "lvio_fusion/common.h"                 // stand-in (oracle/ref_compat): mini-Eigen + Sophus-shaped SE3d
#include "lvio_fusion/ceres/navsat_error.hpp"   // the reference's header, unchanged

using namespace lvio_fusion;
static double urand(unsigned& st) { st = st * 1664525u + 1013904223u; return (double)(st >> 8) / 16777216.0; }

int main() {
    const double yaw_true = -0.45, x_true = 3.5, y_true = 8.25;
    unsigned seed = 7;
    double para[6] = {0, 0, 0, 0, 0, 0};
    ceres::Problem problem;
    problem.AddParameterBlock(para, 1); problem.AddParameterBlock(para + 3, 1); problem.AddParameterBlock(para + 4, 1);
    problem.SetParameterBlockConstant(para + 3); problem.SetParameterBlockConstant(para + 4);
    for (int i = 0; i < 30; ++i) {
        const Vector3d raw(60 * urand(seed) - 30, 60 * urand(seed) - 30, 0.2 * urand(seed));        // navsat point in its own frame
        const Vector3d position(std::cos(yaw_true) * raw.x() - std::sin(yaw_true) * raw.y() + x_true + 0.02 * (urand(seed) - 0.5),
                                std::sin(yaw_true) * raw.x() + std::cos(yaw_true) * raw.y() + y_true + 0.02 * (urand(seed) - 0.5), raw.z());
        ceres::CostFunction* cost_function = NavsatInitError::Create(position, raw, Vector3d(0.01, 0.01, 1.0));
        problem.AddResidualBlock(cost_function, NULL, para, para + 3, para + 4);
    }
    ceres::Solver::Options options;
    options.linear_solver_type = ceres::DENSE_QR;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    problem.SetParameterBlockVariable(para + 3); problem.SetParameterBlockVariable(para + 4);
    ceres::Solve(options, &problem, &summary);
    printf("yaw %.9f x %.9f y %.9f cost %.6e term %d msg %s\n", para[0], para[3], para[4], summary.final_cost, (int)summary.termination_type, summary.message.c_str());
    return (std::fabs(para[0] - yaw_true) < 1e-3 && std::fabs(para[3] - x_true) < 2e-2 && std::fabs(para[4] - y_true) < 2e-2) ? 0 : 1;
}
