// tests/hostcheck/warp_emul.cpp -- TEST HARNESS ONLY.
// Runs lvio_fusion_b200/csrc/lvb_imu_warp.cuh (sqrt_information_warp, the code imu_prepare_kernel executes on the device) on
// the CPU: 32 host threads stand in for the lanes of one warp and move in lock step through a barrier wherever the device code
// converges (__shfl_xor_sync, __syncwarp).  Shared memory is ordinary static / caller-owned memory.  This checks the warp
// algorithm -- the work split over the lanes, the pivot reduction by shuffles, the early return of the Cholesky -- without a GPU;
// it does not check what only hardware can (memory-model visibility between lanes beyond the barriers the code already has).
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace {
class Barrier {
public:
    explicit Barrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const unsigned long gen = gen_;
        if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen_ != gen; });
    }
private:
    std::mutex m_; std::condition_variable cv_; int n_, count_ = 0; unsigned long gen_ = 0;
};
Barrier g_bar(32);
struct Dim { int x; };
thread_local Dim threadIdx;
double g_xd[32]; int g_xi[32];
inline double __shfl_xor_sync(unsigned, double v, int o) { g_xd[threadIdx.x & 31] = v; g_bar.wait(); const double r = g_xd[(threadIdx.x & 31) ^ o]; g_bar.wait(); return r; }
inline int __shfl_xor_sync(unsigned, int v, int o) { g_xi[threadIdx.x & 31] = v; g_bar.wait(); const int r = g_xi[(threadIdx.x & 31) ^ o]; g_bar.wait(); return r; }
inline void __syncwarp() { g_bar.wait(); }
}  // namespace
#define __device__
#define __shared__ static
#define __restrict__
using std::fabs; using std::sqrt;
#include "../../lvio_fusion_b200/csrc/lvb_imu_warp.cuh"

extern "C" int hc_sqrt_information_warp(const double* cov225, double* U225, double prior_a, double prior_g) {
    static double a[225], inv[225];
    int status[32];
    std::vector<std::thread> lanes;
    for (int l = 0; l < 32; ++l) lanes.emplace_back([&, l] { threadIdx.x = l; status[l] = sqrt_information_warp(cov225, U225, a, inv, prior_a, prior_g); });
    for (auto& t : lanes) t.join();
    for (int l = 1; l < 32; ++l) if (status[l] != status[0]) return -100;      // the return value must be warp-uniform
    return status[0];
}
