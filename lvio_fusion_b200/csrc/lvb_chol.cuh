// lvb_chol.cuh -- the two latency-critical pieces of the blocked Cholesky kernels (ba_cholesky_kernel, ba_front_factor_kernel):
// the 32 x 32 diagonal-block factorisation by one warp and the per-row panel solve.
//
// Shared-memory operand: Dt[j * 34 + k] = L[k][j] (the block's factor, column-major with a 16-byte aligned stride), invd[j] = 1 / L[j][j].
#pragma once

// One warp, lane = row of the block, a[r] = entry (lane, c0 + r) of the block (lower triangle valid), bn = rows/columns in use.
// The 32 columns go in 4 groups of 8.  Inside a group a finished column is applied at once only to the group's own columns (<= 7
// FMAs per lane, operands broadcast from Dt with <= 4 LDS.128); the columns right of the group receive the group's rank-8 update
// in one sweep afterwards.  Per column the chain is mul -> FMA -> shuffle -> rsqrt, and a single warp issues it at ~0.25 IPC, so
// the instruction count per column is what matters: entries above the diagonal are don't-care and are updated without per-lane
// predicates (garbage stays garbage, nothing valid ever reads it), and nothing is stored to global memory here -- the caller
// copies the finished block (Dt, invd) to its place with all threads at the start of the next panel phase.
__device__ __forceinline__ int chol_diag32(double (&a)[32], const int lane, double* __restrict__ Dt, double* __restrict__ invd) {
    int bad = 0;
    double d0 = __shfl_sync(0xffffffffu, a[0], 0);
    if (!(d0 > 0.0)) { bad = 1; d0 = 1.0; }
    double inv = rsqrt(d0);
#pragma unroll 1
    for (int c0 = 0; c0 < 32; c0 += 8) {
        const int rel = lane - c0;                    // register index of this lane's diagonal entry
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = c0 + jj;
            a[jj] *= inv;                             // l_ij for lanes >= j (lane j: sqrt(d_jj)); don't-care above the diagonal
            Dt[j * 34 + lane] = a[jj];
            if (rel == jj) invd[j] = inv;
            // next pivot inside the group: lane j + 1 owns everything its diagonal entry still needs
            double inv_next = 1.0;
            if (jj < 7) {
                double dn = __shfl_sync(0xffffffffu, a[jj + 1] - a[jj] * a[jj], j + 1);
                if (!(dn > 0.0)) { bad = 1; dn = 1.0; }
                inv_next = rsqrt(dn);
            }
            __syncwarp();
            const double2* bp = reinterpret_cast<const double2*>(Dt + j * 34 + c0);      // L[c0 + r][j], r = 0..7
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (2 * p + 1 > jj) {
                    const double2 v = bp[p];
                    if (2 * p > jj) a[2 * p] -= a[jj] * v.x;
                    a[2 * p + 1] -= a[jj] * v.y;
                }
            }
            inv = inv_next;
        }
        if (c0 < 24) {
            // rank-8 update of the columns right of the group: a[r] -= sum_k L[lane][c0 + k] * L[c0 + r][c0 + k].  Register 8 (the next
            // pivot's column) first, so that its shuffle + rsqrt overlap the rest of the sweep.
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double2 v = *reinterpret_cast<const double2*>(Dt + (c0 + k) * 34 + c0 + 8);
                a[8] -= a[k] * v.x; a[9] -= a[k] * v.y;
            }
            double dn = __shfl_sync(0xffffffffu, a[8], c0 + 8);
            if (!(dn > 0.0)) { bad = 1; dn = 1.0; }
            inv = rsqrt(dn);
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
