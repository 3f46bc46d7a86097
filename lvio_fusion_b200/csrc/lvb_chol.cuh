// lvb_chol.cuh -- the two latency-critical pieces of the blocked Cholesky kernels (ba_cholesky_kernel, ba_front_factor_kernel):
// the 32 x 32 diagonal-block factorisation by one warp and the per-row panel solve.
//
// Shared-memory operand: Dt[j * 34 + k] = L[k][j] (the block's factor, column-major with a 16-byte aligned stride), invd[j] = 1 / L[j][j].
#pragma once

// 1 / sqrt(x) for a positive normal x without the library routine's special-case branch: a single warp is an in-order machine, and
// that branch (plus the call behind it) fences the scheduler, so nothing independent could be interleaved into the pivot chain's
// latency.  MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-20) + two coupled Newton (Goldschmidt) steps: 2^-40, then below 2^-52.
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    h = fma(h, r, h);
    return h + h;
}

// One warp, lane = row of the block, a[r] = entry (lane, c0 + r) of the block (lower triangle valid).
// The 32 columns go in 4 groups of 8.  Inside a group a finished column is applied at once only to the group's own columns: its
// entries L[c0 + r][j] come from the owning lanes by shuffle (no shared-memory round trip, no warp barrier in the chain); the
// columns right of the group receive the group's rank-8 update in one sweep afterwards, operands broadcast from Dt as LDS.128.
// Per column the dependent chain is mul -> shuffle -> FMA -> shuffle -> rsqrt (~110 cycles); everything else is independent work
// the scheduler can place into its shadow because the loop body is branch-free.  Entries above the diagonal are don't-care and are
// updated without per-lane predicates (garbage stays garbage, nothing valid ever reads it), and nothing is stored to global memory
// here -- the caller copies the finished block (Dt, invd) to its place with all threads at the start of the next panel phase.
__device__ __forceinline__ int chol_diag32(double (&a)[32], const int lane, double* __restrict__ Dt, double* __restrict__ invd) {
    int bad = 0;
    double d0 = __shfl_sync(0xffffffffu, a[0], 0);
    if (!(d0 > 0.0)) { bad = 1; d0 = 1.0; }
    double inv = rsqrt_pos(d0);
#pragma unroll 1
    for (int c0 = 0; c0 < 32; c0 += 8) {
        const int rel = lane - c0;                    // register index of this lane's diagonal entry
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = c0 + jj;
            a[jj] *= inv;                             // l_ij for lanes >= j (lane j: sqrt(d_jj)); don't-care above the diagonal
            Dt[j * 34 + lane] = a[jj];
            if (rel == jj) invd[j] = inv;
#pragma unroll
            for (int r = jj + 1; r < 8; ++r) {        // r = jj + 1 first: it completes the next pivot
                const double v = __shfl_sync(0xffffffffu, a[jj], c0 + r);
                a[r] -= a[jj] * v;
                if (r == jj + 1) {
                    double dn = __shfl_sync(0xffffffffu, a[jj + 1], j + 1);
                    if (!(dn > 0.0)) { bad = 1; dn = 1.0; }
                    inv = rsqrt_pos(dn);
                }
            }
        }
        if (c0 < 24) {
            __syncwarp();                             // the group's columns are in Dt
            // rank-8 update of the columns right of the group: a[r] -= sum_k L[lane][c0 + k] * L[c0 + r][c0 + k].  Register 8 (the next
            // pivot's column) first, so that its shuffle + rsqrt overlap the rest of the sweep.
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double2 v = *reinterpret_cast<const double2*>(Dt + (c0 + k) * 34 + c0 + 8);
                a[8] -= a[k] * v.x; a[9] -= a[k] * v.y;
            }
            double dn = __shfl_sync(0xffffffffu, a[8], c0 + 8);
            if (!(dn > 0.0)) { bad = 1; dn = 1.0; }
            inv = rsqrt_pos(dn);
#pragma unroll
            for (int r = 10; r < 32; r += 2) {
                if (c0 + r < 32) {                    // warp-uniform
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const double2 v = *reinterpret_cast<const double2*>(Dt + (c0 + k) * 34 + c0 + r);
                        a[r] -= a[k] * v.x; a[r + 1] -= a[k] * v.y;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 24; ++r) a[r] = a[r + 8];
#pragma unroll
        for (int r = 24; r < 32; ++r) a[r] = 0.0;
    }
    __syncwarp();
    return bad;
}

// 1 / sqrt(x), seed + ONE third-order step  y (1 + e/2 + 3 e^2/8),  e = 1 - x y^2: the seed is good to ~2^-20 (tools/lat_probe.cu prints
// the measured worst case), the truncation term 5/16 e^3 is then below 2^-58, and the dependent chain is MUFU + 4 FP64 operations
// instead of MUFU + 6.
__device__ __forceinline__ double rsqrt_pos3(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(e, 0.375, 0.5), ye = y * e;
    return fma(ye, p, y);
}

// Same contract as chol_diag32, two columns per pivot step.  The 2 x 2 pivot block [[p, q], [q, c]] is factored in closed form:
// L11 = sqrt(p), L21 = q / L11, L22 = sqrt(det / p) with det = p c - q^2, so the two reciprocal square roots 1 / sqrt(p) and 1 / sqrt(det)
// are independent and run side by side -- one MUFU + refinement latency per TWO columns -- and the broadcast / update round trip
// (shuffle -> FMA -> shuffle) is also paid once per pair.
__device__ __forceinline__ int chol_diag32_pair(double (&a)[32], const int lane, double* __restrict__ Dt, double* __restrict__ invd) {
    int bad = 0;
#pragma unroll 1
    for (int c0 = 0; c0 < 32; c0 += 8) {
#pragma unroll
        for (int jj = 0; jj < 8; jj += 2) {
            const int j = c0 + jj;
            double p = __shfl_sync(0xffffffffu, a[jj], j);
            const double q = __shfl_sync(0xffffffffu, a[jj], j + 1);
            const double c = __shfl_sync(0xffffffffu, a[jj + 1], j + 1);
            double det = fma(p, c, -q * q);
            if (!(p > 0.0) || !(det > 0.0)) { bad = 1; p = 1.0; det = 1.0; }
            const double i1 = rsqrt_pos3(p), id = rsqrt_pos3(det);
            const double l11 = p * i1, l21 = q * i1;
            const double i2 = id * l11;
            const double l1 = a[jj] * i1;
            const double l2 = fma(-l1, l21, a[jj + 1]) * i2;
            a[jj] = l1; a[jj + 1] = l2;
            Dt[j * 34 + lane] = l1; Dt[(j + 1) * 34 + lane] = l2;
            if (lane == j) { invd[j] = i1; invd[j + 1] = i2; }
#pragma unroll
            for (int r = jj + 2; r < 8; ++r) {
                const double v1 = __shfl_sync(0xffffffffu, l1, c0 + r), v2 = __shfl_sync(0xffffffffu, l2, c0 + r);
                a[r] = fma(-l2, v2, fma(-l1, v1, a[r]));
            }
        }
        if (c0 < 24) {
            __syncwarp();                             // the group's columns are in Dt
            // rank-8 update of the columns right of the group; registers 8 and 9 (the next pivot pair) first
#pragma unroll
            for (int r = 8; r < 32; r += 2) {
                if (c0 + r < 32) {                    // warp-uniform
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const double2 v = *reinterpret_cast<const double2*>(Dt + (c0 + k) * 34 + c0 + r);
                        a[r] -= a[k] * v.x; a[r + 1] -= a[k] * v.y;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 24; ++r) a[r] = a[r + 8];
#pragma unroll
        for (int r = 24; r < 32; ++r) a[r] = 0.0;
    }
    __syncwarp();
    return bad;
}

// One panel row: x L^T = a for the 32 columns of the block (right-looking, no divisions), operands broadcast from Dt as LDS.128.
__device__ __forceinline__ void chol_panel_row(double (&a)[32], const double* __restrict__ Dt, const double* __restrict__ invd) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        a[j] *= invd[j];
        const double2* bp = reinterpret_cast<const double2*>(Dt + j * 34);
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            if (2 * p + 1 > j) {
                const double2 v = bp[p];
                if (2 * p > j) a[2 * p] -= a[j] * v.x;
                a[2 * p + 1] -= a[j] * v.y;
            }
        }
    }
}
