// oracle/ref_time_harness.cpp -- TEST / BENCH INFRASTRUCTURE ONLY.
// CPU counterpart of the roofline kernel (ba_eval_two_frame_kernel): the REFERENCE's own TwoFrameReprojectionError functor
// (include/lvio_fusion/ceres/visual_error.hpp:78-107, compiled where it lies) evaluated the way Ceres' AutoDiffCostFunction
// evaluates a residual block -- operator() instantiated on 15-wide dual numbers: residual + the full 2 x 15 Jacobian -- over
// seeded KITTI-shaped blocks, one thread.  Prints  blocks  seconds  blocks_per_s .  Built by `make -C oracle ref` into
// oracle/_ref/ref_time where /root/reference is mounted; bench.py runs the prebuilt binary beside the roofline leg.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lvio_fusion/ceres/visual_error.hpp"

namespace lvio_fusion {
std::vector<Camera::Ptr> Camera::devices_;
double Camera::baseline = 1;
}
using namespace lvio_fusion;
using oracle::Dual;

static unsigned g_seed = 12345u;
static double urand() { g_seed = g_seed * 1664525u + 1013904223u; return (double)(g_seed >> 8) / 16777216.0; }

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 20000;
    const double budget = argc > 2 ? std::atof(argv[2]) : 2.0;            // seconds
    const double e0[7] = {0, 0, 0, 1, 0, 0, 0}, e1[7] = {0, 0, 0, 1, 0.537, 0, 0};
    Camera::Create(718.856, 718.856, 607.1928, 185.2157, SE3d(e0));
    Camera::Create(718.856, 718.856, 607.1928, 185.2157, SE3d(e1));
    Camera::Ptr left = Camera::Get(0), right = Camera::Get(1);
    std::vector<TwoFrameReprojectionError> f;
    std::vector<double> x((size_t)n * 15);
    f.reserve(n);
    for (int i = 0; i < n; ++i) {
        f.emplace_back(Vector2d(1241 * urand(), 376 * urand()), Vector2d(1241 * urand(), 376 * urand()), left, right, 71.9);
        double* p = &x[(size_t)i * 15];
        p[0] = 0.02 + 0.2 * urand();                                              // inverse depth
        for (int b = 0; b < 2; ++b) {                                             // two keyframe poses, small rotations
            double* T = p + 1 + 7 * b;
            T[0] = 0.02 * (urand() - 0.5); T[1] = 0.02 * (urand() - 0.5); T[2] = 0.05 * (urand() - 0.5); T[3] = 1.0;
            T[4] = 2.0 * b + 0.1 * urand(); T[5] = 0.1 * urand(); T[6] = 0.05 * urand();
        }
    }
    double sink = 0;
    size_t blocks = 0;
    const auto t0 = std::chrono::steady_clock::now();
    double dt = 0;
    do {
        for (int i = 0; i < n; ++i) {
            Dual<15> X[15], Y[2];
            const double* p = &x[(size_t)i * 15];
            for (int k = 0; k < 15; ++k) X[k] = Dual<15>::seed(p[k], k);
            f[i](X, X + 1, X + 8, Y);
            sink += Y[0].v + Y[1].d[3];
        }
        blocks += n;
        dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } while (dt < budget);
    printf("%zu %.6f %.1f %.3e\n", blocks, dt, blocks / dt, sink);
    return 0;
}
