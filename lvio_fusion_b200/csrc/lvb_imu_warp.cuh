// lvio_fusion_b200/csrc/lvb_imu_warp.cuh -- the warp-cooperative whitening of one IMU factor (used by imu_prepare_kernel, ba.cu).
// U = LLT(cov^-1 [bias blocks overwritten by the priors]).matrixL()^T  (imu_error.hpp:32,147-150), one warp per factor:
// same arithmetic per element as lvb_math.cuh::sqrt_information (partial-pivot LU inverse, then the unblocked lower Cholesky with
// Eigen's early return on a non-positive pivot; the oracle's order), the independent elements of every step spread over the
// lanes.  `a` and `inv` are 225-double scratch arrays in shared memory owned by the warp.  Returns 0, or 1 (singular) / 2 (NaN
// pivot), uniformly across the warp.
// Kept in its own header so that tests/hostcheck/warp_emul.cpp can run it on the CPU with 32 lock-step host threads standing in
// for the lanes (the shuffles and __syncwarp are the only warp primitives it uses).
#pragma once

__device__ inline int sqrt_information_warp(const double* __restrict__ cov, double* __restrict__ U, double* a, double* inv, double prior_a, double prior_g) {
    const int n = 15, lane = threadIdx.x & 31;
    __shared__ int s_piv[4][16];
    int* piv = s_piv[(threadIdx.x >> 5) & 3];
    for (int e = lane; e < 225; e += 32) a[e] = cov[e];
    if (lane < n) piv[lane] = lane;
    __syncwarp();
    for (int k = 0; k < n; ++k) {
        // pivot: largest |a[i][k]|, i >= k, first index on ties
        double bv = (lane >= k && lane < n) ? fabs(a[lane * n + k]) : -1.0;
        int best = lane;
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int ob = __shfl_xor_sync(0xffffffffu, best, o);
            if (ov > bv || (ov == bv && ob < best)) { bv = ov; best = ob; }
        }
        if (bv == 0.0) return 1;
        if (best != k) {
            if (lane < n) { const double t = a[k * n + lane]; a[k * n + lane] = a[best * n + lane]; a[best * n + lane] = t; }
            if (lane == 0) { const int t = piv[k]; piv[k] = piv[best]; piv[best] = t; }
        }
        __syncwarp();
        if (lane > k && lane < n) a[lane * n + k] /= a[k * n + k];
        __syncwarp();
        const int m = n - 1 - k;
        for (int e = lane; e < m * m; e += 32) { const int i = k + 1 + e / m, j = k + 1 + e % m; a[i * n + j] -= a[i * n + k] * a[k * n + j]; }
        __syncwarp();
    }
    if (lane < n) {                       // column `lane` of the inverse: forward / backward substitution
        const int c = lane;
        double y[15];
#pragma unroll
        for (int i = 0; i < 15; ++i) {
            double sacc = (piv[i] == c) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 15; ++k) if (k < i) sacc -= a[i * n + k] * y[k];
            y[i] = sacc;
        }
#pragma unroll
        for (int i = 14; i >= 0; --i) {
            double sacc = y[i];
#pragma unroll
            for (int k = 0; k < 15; ++k) if (k > i) sacc -= a[i * n + k] * inv[k * n + c];
            inv[i * n + c] = sacc / a[i * n + i];
        }
    }
    __syncwarp();
    if (prior_a >= 0.0 && prior_g >= 0.0 && lane < 9) {      // ImuInitError (imu_error.hpp:147-149)
        const int i = lane / 3, j = lane % 3;
        inv[(9 + i) * n + 9 + j] = (i == j) ? prior_a : 0.0; inv[(12 + i) * n + 12 + j] = (i == j) ? prior_g : 0.0;
    }
    for (int e = lane; e < 225; e += 32) a[e] = 0.0;          // a := L
    __syncwarp();
    for (int j = 0; j < n; ++j) {
        double d = inv[j * n + j];
        for (int k = 0; k < j; ++k) d -= a[j * n + k] * a[j * n + k];
        if (d != d) return 2;
        if (d <= 0.0) {      // warp-uniform.  Eigen's LLT stops here; matrixL() shows the untouched lower triangle from column j on
            for (int e = lane; e < 225; e += 32) { const int i = e / n, c = e - n * i; if (c >= j && i >= c) a[e] = inv[e]; }
            __syncwarp();
            break;
        }
        const double ljj = sqrt(d);
        if (lane == 0) a[j * n + j] = ljj;
        if (lane > j && lane < n) { double sacc = inv[lane * n + j]; for (int k = 0; k < j; ++k) sacc -= a[lane * n + k] * a[j * n + k]; a[lane * n + j] = sacc / ljj; }
        __syncwarp();
    }
    for (int e = lane; e < 225; e += 32) { const int i = e / n, j = e - n * i; U[e] = a[j * n + i]; }
    return 0;
}
