// icp.cu -- lidar scan-to-map registration on the device: the kernels behind
// pcl::KdTreeFLANN + FeatureAssociation::ScanToMapWith{Ground,Segmented} + the two DENSE_QR solves
// of Mapping::Optimize (/root/reference/src/lvio_fusion/src/association.cpp:270-384,
// src/mapping.cpp:139-191).
//
//   K7  voxel hash build : key = ix + gx*(iy + gy*iz) over the cloud's bounding box (the PCL VoxelGrid
//       keying), counting sort (histogram -> exclusive scan -> scatter) = one radix pass with radix #cells
//   K8  transform + exact 3-NN + gate : float32, no fused multiply-add (__f*_rn), ascending (d2, index)
//   K9  point-to-plane residual/Jacobian + 3x3 normal equations reduce + LM on three scalars
//
// HBM layout: map cloud re-packed as float4 (x, y, z, original index bits) sorted by voxel key (16 B/pt);
// cell_start int32[#cells+1]; scan cloud read in place from the caller's record stride after upload;
// per-query association: accepted u8, pa float[3][K], normal double[3][K] (SoA planes).
#include <float.h>
#include <math.h>
#include <algorithm>
#include <chrono>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "lvb_internal.cuh"
#include "lvb_math.cuh"
#include "lvb_cloud.cuh"

using namespace lvb;

namespace {

enum { ITPB = 128, MAX_CELLS = 1 << 26 };

struct Grid {
    float minx, miny, minz, inv_cell, cell;   // cell = voxel edge = (search radius * 1.0001) / ring
    int gx, gy, gz;
    int ring;                                 // voxels per search radius: neighbours within the radius lie within `ring` voxels
};

struct IcpState {
    LmState lm;
    int mode;
    double Twc1[7];
    double rpyxyz[6];
    double x[3], cand[3], delta[3], lam[3], scale[3], grad[3];
    double prior_w, prior_target[3], huber_a, weight;
    double acc[16];      // consumed system: H (9, row-major), g (3), cost, n_accepted
    double part[16];     // per-pass partial sums written by the linearize kernels (all-reduced when world > 1): [0..13] as acc, [14] candidate cost
};

struct IcpDev {
    const float4* map;       // sorted by cell
    const int* cell_start;
    Grid g;
    int P;
    const unsigned char* scan;   // device copy of the caller's records
    int K, stride;
    float tf[7];                 // frame pose cast to float (association.cpp:287)
    float max_d2;
    double thr;
    unsigned char* accepted;
    float* pa;                   // 3 planes
    double* nrm;                 // 3 planes
    IcpState* st;
    int rank, world;
    const int* order;            // query visiting order (spatially sorted) or nullptr
    const int* coarse;           // occupancy of the 4 x 4 x 4-voxel blocks of the map grid (1: holds points)
    int cbx, cby, cbz;           // its dimensions
};

__device__ __forceinline__ float3 load_xyz(const unsigned char* base, int i, int stride) {
    const float* f = reinterpret_cast<const float*>(base + (size_t)i * stride);
    return make_float3(f[0], f[1], f[2]);
}

__global__ void icp_bbox_kernel(const unsigned char* pts, int n, int stride, int* bbox /*6: min xyz, max xyz (ordered ints)*/) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float3 p = load_xyz(pts, i, stride);
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) { mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o)); mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(&bbox[a], f2ord(mn[a])); atomicMax(&bbox[3 + a], f2ord(mx[a])); }
    }
}

__device__ __forceinline__ int cell_coord(float v, float mn, float inv_cell) { return (int)floorf((v - mn) * inv_cell); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void icp_count_kernel(const unsigned char* pts, int n, int stride, Grid g, int* cell_of, int* counts, int* coarse, int cbx, int cby) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 p = load_xyz(pts, i, stride);
    const int ix = clampi(cell_coord(p.x, g.minx, g.inv_cell), 0, g.gx - 1);
    const int iy = clampi(cell_coord(p.y, g.miny, g.inv_cell), 0, g.gy - 1);
    const int iz = clampi(cell_coord(p.z, g.minz, g.inv_cell), 0, g.gz - 1);
    const int c = ix + g.gx * (iy + g.gy * iz);
    cell_of[i] = c;
    atomicAdd(&counts[c], 1);
    coarse[(ix >> 2) + cbx * ((iy >> 2) + cby * (iz >> 2))] = 1;       // benign race: every writer stores 1
}

__global__ void icp_scatter_kernel(const unsigned char* pts, int n, int stride, const int* cell_of, const int* cell_start, int* fill, float4* sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 p = load_xyz(pts, i, stride);
    const int c = cell_of[i];
    const int pos = cell_start[c] + atomicAdd(&fill[c], 1);
    sorted[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
}

// Mapping::MergeScan (mapping.cpp:193-203): world cloud = float32 SE3 transform of the robot-frame cloud; the other
// fields of each record (intensity, padding) are carried over unchanged.
__global__ void icp_transform_cloud_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int n, int stride, const float* __restrict__ tf7) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float tf[7];
    for (int k = 0; k < 7; ++k) tf[k] = tf7[k];
    const float* src = reinterpret_cast<const float*>(in + (size_t)i * stride);
    float* dst = reinterpret_cast<float*>(out + (size_t)i * stride);
    const float3 q = transform_f32(tf, make_float3(src[0], src[1], src[2]));
    dst[0] = q.x; dst[1] = q.y; dst[2] = q.z;
    for (int k = 3; k < stride / 4; ++k) dst[k] = src[k];
}

// Mapping::ToWorld (mapping.cpp:205-220) into the resident store: records of any stride -> packed (x, y, z, intensity) in the world frame
__global__ void icp_to_world_kernel(const unsigned char* __restrict__ in, float4* __restrict__ out, int n, int stride, const float* __restrict__ tf7) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float tf[7];
    for (int k = 0; k < 7; ++k) tf[k] = tf7[k];
    const float* src = reinterpret_cast<const float*>(in + (size_t)i * stride);
    const float3 q = transform_f32(tf, make_float3(src[0], src[1], src[2]));
    out[i] = make_float4(q.x, q.y, q.z, stride >= 16 ? src[3] : 0.0f);
}

struct Best3 { float d[3]; int i[3]; int p[3]; };   // d2, original index, position in the sorted array
__device__ __forceinline__ void best_insert(Best3& b, float d, int i, int p) {
    if (d > b.d[2] || (d == b.d[2] && i > b.i[2])) return;
    if (d < b.d[1] || (d == b.d[1] && i < b.i[1])) {
        b.d[2] = b.d[1]; b.i[2] = b.i[1]; b.p[2] = b.p[1];
        if (d < b.d[0] || (d == b.d[0] && i < b.i[0])) { b.d[1] = b.d[0]; b.i[1] = b.i[0]; b.p[1] = b.p[0]; b.d[0] = d; b.i[0] = i; b.p[0] = p; }
        else { b.d[1] = d; b.i[1] = i; b.p[1] = p; }
    } else { b.d[2] = d; b.i[2] = i; b.p[2] = p; }
}

// exact 3-NN of q among map points with d2 <= max_d2 (max_d2 <= radius^2).  Voxels are visited in Chebyshev
// rings around the query's voxel; a ring r >= 1 cannot hold a point closer than (r-1)*cell, so the search stops
// as soon as the third-best distance is strictly below that bound (strict: ties are broken by index).
struct Occ { unsigned mask; int bx, by, bz; };        // occupancy bits of the 3 x 3 x 3 coarse blocks around the query's block (bx, by, bz)
__device__ __forceinline__ void scan_voxel(const IcpDev& d, const Grid& g, float3 q, int ix, int iy, int iz, Best3& b, const Occ& occ) {
    // an empty coarse block holds no points: skip without touching memory.  A query without neighbours otherwise walks all 729 voxels
    // of the radius at two dependent L2 round trips each (~0.5 ms for ONE thread), and the kernel lasts as long as its slowest thread.
    const int bit = ((iz >> 2) - occ.bz + 1) * 9 + ((iy >> 2) - occ.by + 1) * 3 + ((ix >> 2) - occ.bx + 1);
    if (!((occ.mask >> bit) & 1u)) return;
    if (ix < 0 || iy < 0 || iz < 0 || ix >= g.gx || iy >= g.gy || iz >= g.gz) return;
    const int c = ix + g.gx * (iy + g.gy * iz);
    const int s = d.cell_start[c], e = d.cell_start[c + 1];
    if (s == e) return;
    const float bx0 = g.minx + ix * g.cell, by0 = g.miny + iy * g.cell, bz0 = g.minz + iz * g.cell;
    const float ex = fmaxf(fmaxf(bx0 - q.x, q.x - (bx0 + g.cell)), 0.0f);
    const float ey = fmaxf(fmaxf(by0 - q.y, q.y - (by0 + g.cell)), 0.0f);
    const float ez = fmaxf(fmaxf(bz0 - q.z, q.z - (bz0 + g.cell)), 0.0f);
    const float bd = (ex * ex + ey * ey + ez * ez) * 0.99f - 1e-6f;     // conservative box distance
    if (bd > b.d[2] || bd > d.max_d2) return;
    for (int j = s; j < e; ++j) {
        const float4 m = __ldg(&d.map[j]);
        const float dx = __fsub_rn(m.x, q.x), dy = __fsub_rn(m.y, q.y), dz = __fsub_rn(m.z, q.z);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        best_insert(b, d2, __float_as_int(m.w), j);
    }
}

__device__ __forceinline__ Best3 knn3_query(const IcpDev& d, float3 q) {
    Best3 b;
    b.d[0] = b.d[1] = b.d[2] = INFINITY; b.i[0] = b.i[1] = b.i[2] = 0x7fffffff; b.p[0] = b.p[1] = b.p[2] = -1;
    const Grid& g = d.g;
    const int cx = cell_coord(q.x, g.minx, g.inv_cell), cy = cell_coord(q.y, g.miny, g.inv_cell), cz = cell_coord(q.z, g.minz, g.inv_cell);
    Occ occ; occ.mask = 0u; occ.bx = cx >> 2; occ.by = cy >> 2; occ.bz = cz >> 2;       // (ring <= 4 voxels: every probed voxel lies in these 27 blocks)
#pragma unroll
    for (int k = 0; k < 27; ++k) {                  // 27 independent loads, one round trip
        const int x = occ.bx + k % 3 - 1, y = occ.by + (k / 3) % 3 - 1, z = occ.bz + k / 9 - 1;
        const bool in = x >= 0 && y >= 0 && z >= 0 && x < d.cbx && y < d.cby && z < d.cbz;
        if (in && __ldg(&d.coarse[x + d.cbx * (y + d.cby * z)])) occ.mask |= 1u << k;
    }
    scan_voxel(d, g, q, cx, cy, cz, b, occ);
    for (int r = 1; r <= g.ring; ++r) {
        const float lb = (r - 1) * g.cell;
        const float lb2 = lb * lb * 0.99f;
        if (b.d[2] < lb2 || lb2 > d.max_d2) break;
        // the six faces of the ring-r shell, each voxel exactly once
        for (int dz = -r; dz <= r; ++dz) {
            const bool zface = (dz == -r || dz == r);
            for (int dy = -r; dy <= r; ++dy) {
                const bool yface = (dy == -r || dy == r);
                if (zface || yface) { for (int dx = -r; dx <= r; ++dx) scan_voxel(d, g, q, cx + dx, cy + dy, cz + dz, b, occ); }
                else { scan_voxel(d, g, q, cx - r, cy + dy, cz + dz, b, occ); scan_voxel(d, g, q, cx + r, cy + dy, cz + dz, b, occ); }
            }
        }
    }
    for (int j = 0; j < 3; ++j) if (!(b.d[j] <= d.max_d2)) { b.d[j] = INFINITY; b.i[j] = -1; b.p[j] = -1; }
    return b;
}

// Coarse spatial key of a query (blocks of 4x4x4 voxels) so that the lanes of a warp walk the same voxel lists and
// their 16-byte map loads hit the same L1 lines instead of 32 different L2 sectors.
// Measured alternatives that lost against this per-query ring walk (round 2, 120 k queries vs 1 M points, 0.50 ms): scanning
// x-major voxel ROWS as contiguous runs (fewer probes, but no per-voxel culling: 0.78 ms); staging a coarse block's 3 x 3 x 3
// neighbourhood in shared memory per CTA and searching there (ring search on the staged copy 1.4 - 1.8 ms, brute force 4.4 ms):
// a block holds ~25 queries that look at ~50 points each, while its neighbourhood is ~3 000 points -- the staging reads 60x more
// map than the queries need; four lanes per query (a ring's voxel rows dealt round-robin, butterfly merge of the best three after
// every ring): 0.57 ms -- the lanes prune with their own looser bounds and scan more voxels than the chain they shorten; runs only for
// the face rows of the outer rings (r >= 2): 0.57 ms.  The kernel takes ~0.47 - 0.50 ms whether a rank holds 120 k, 60 k or 30 k of
// the queries (round-2 runs at 1 / 2 / 4 GPUs): its duration is set by its slowest warps, not by the query count.
__global__ void icp_query_key_kernel(IcpDev d, int shift, int cgx, int cgy, int cgz, int* __restrict__ key, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.K) return;
    const float3 q = transform_f32(d.tf, load_xyz(d.scan, i, d.stride));
    const Grid& g = d.g;
    const int ix = clampi(cell_coord(q.x, g.minx, g.inv_cell), 0, g.gx - 1) >> shift;
    const int iy = clampi(cell_coord(q.y, g.miny, g.inv_cell), 0, g.gy - 1) >> shift;
    const int iz = clampi(cell_coord(q.z, g.minz, g.inv_cell), 0, g.gz - 1) >> shift;
    const int c = ix + cgx * (iy + cgy * iz);
    key[i] = c;
    atomicAdd(&counts[c], 1);
}
__global__ void icp_query_scatter_kernel(int n, const int* __restrict__ key, const int* __restrict__ start, int* __restrict__ fill, int* __restrict__ order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = key[i];
    order[start[c] + atomicAdd(&fill[c], 1)] = i;
}

__global__ void __launch_bounds__(ITPB) icp_knn_kernel(IcpDev d, int* __restrict__ idx_out, float* __restrict__ d2_out) {
    const int t = blockIdx.x * ITPB + threadIdx.x;
    if (t >= d.K) return;
    const int i = d.order ? d.order[t] : t;
    const float3 q = transform_f32(d.tf, load_xyz(d.scan, i, d.stride));
    const Best3 b = knn3_query(d, q);
    for (int j = 0; j < 3; ++j) { idx_out[3 * i + j] = b.i[j]; d2_out[3 * i + j] = b.d[j]; }
}

// K8: transform + 3-NN + gate (association.cpp:296-300) + plane constants (lidar_error.hpp:13-18)
__global__ void __launch_bounds__(ITPB) icp_associate_kernel(IcpDev d) {
    const int t = blockIdx.x * ITPB + threadIdx.x;
    if (t >= d.K) return;
    const int i = d.order ? d.order[t] : t;
    const float3 q = transform_f32(d.tf, load_xyz(d.scan, i, d.stride));
    const Best3 b = knn3_query(d, q);
    int ok = 1;
    for (int j = 0; j < 3; ++j) if (!(b.i[j] >= 0 && b.i[j] < d.P && (double)b.d[j] < d.thr)) ok = 0;
    d.accepted[i] = (unsigned char)ok;
    if (!ok) return;
    const float4 a = d.map[b.p[0]], bb = d.map[b.p[1]], c = d.map[b.p[2]];
    const V3 n = plane_normal(v3(a.x, a.y, a.z), v3(bb.x, bb.y, bb.z), v3(c.x, c.y, c.z));
    d.pa[i] = a.x; d.pa[d.K + i] = a.y; d.pa[2 * d.K + i] = a.z;
    d.nrm[i] = n.x; d.nrm[d.K + i] = n.y; d.nrm[2 * (size_t)d.K + i] = n.z;
}

__device__ __forceinline__ void icp_substitute(const IcpState& s, const double* x, double* e) {
    for (int k = 0; k < 6; ++k) e[k] = s.rpyxyz[k];
    if (s.mode == 0) { e[1] = x[0]; e[2] = x[1]; e[5] = x[2]; } else { e[0] = x[0]; e[3] = x[1]; e[4] = x[2]; }
}

// parity entry: raw residual / Jacobian per scan point at st->x
__global__ void __launch_bounds__(ITPB) icp_eval_kernel(IcpDev d, double* __restrict__ r_out, double* __restrict__ J_out) {
    __shared__ IcpFrame s_f;
    if (threadIdx.x == 0) { double e[6]; icp_substitute(*d.st, d.st->x, e); s_f = icp_frame(d.st->mode, d.st->Twc1, e); }
    __syncthreads();
    const int i = blockIdx.x * ITPB + threadIdx.x;
    if (i >= d.K) return;
    double r = 0.0, J[3] = {0, 0, 0};
    if (d.accepted[i]) {
        const float3 p = load_xyz(d.scan, i, d.stride);
        r = icp_point(s_f, v3(p.x, p.y, p.z), v3(d.pa[i], d.pa[d.K + i], d.pa[2 * d.K + i]),
                      v3(d.nrm[i], d.nrm[d.K + i], d.nrm[2 * (size_t)d.K + i]), d.st->weight, J);
    }
    r_out[i] = r; J_out[3 * i] = J[0]; J_out[3 * i + 1] = J[1]; J_out[3 * i + 2] = J[2];
}

// K9: residual / Jacobian per accepted point, Huber corrector, block reduce of J^T J (3x3), J^T r, cost.
// MODE 0: at x (skipped unless need_linearize) ; MODE 1: cost at the candidate only.
template <int MODE>
__global__ void __launch_bounds__(ITPB) icp_linearize_kernel(IcpDev d) {
    __shared__ IcpFrame s_f;
    __shared__ double s_red[ITPB / 32][12];
    IcpState* st = d.st;
    if (st->lm.done) return;
    if (MODE == 0 && !st->lm.need_linearize) return;
    if (threadIdx.x == 0) { double e[6]; icp_substitute(*st, MODE == 0 ? st->x : st->cand, e); s_f = icp_frame(st->mode, st->Twc1, e); }
    __syncthreads();
    double a[11];
    for (int k = 0; k < 11; ++k) a[k] = 0.0;     // H00 H10 H11 H20 H21 H22 g0 g1 g2 cost count
    for (int i = blockIdx.x * ITPB + threadIdx.x; i < d.K; i += gridDim.x * ITPB) {
        if (!d.accepted[i]) continue;
        const float3 p = load_xyz(d.scan, i, d.stride);
        double J[3];
        double r = icp_point(s_f, v3(p.x, p.y, p.z), v3(d.pa[i], d.pa[d.K + i], d.pa[2 * d.K + i]),
                             v3(d.nrm[i], d.nrm[d.K + i], d.nrm[2 * (size_t)d.K + i]), st->weight, MODE == 0 ? J : nullptr);
        double rho_v, sr;
        huber(st->huber_a, r * r, &rho_v, &sr);
        a[9] += 0.5 * rho_v; a[10] += 1.0;
        if (MODE == 0) {
            r *= sr; J[0] *= sr; J[1] *= sr; J[2] *= sr;
            a[0] += J[0] * J[0]; a[1] += J[1] * J[0]; a[2] += J[1] * J[1]; a[3] += J[2] * J[0]; a[4] += J[2] * J[1]; a[5] += J[2] * J[2];
            a[6] += J[0] * r; a[7] += J[1] * r; a[8] += J[2] * r;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < 11; ++k) { double v = a[k]; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); if (lane == 0) s_red[warp][k] = v; }
    __syncthreads();
    if (threadIdx.x < 11) {
        double v = 0.0; for (int w = 0; w < ITPB / 32; ++w) v += s_red[w][threadIdx.x];
        if (MODE == 0) {
            static const int map[11] = {0, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13};
            if (v != 0.0) atomicAdd(&st->part[map[threadIdx.x]], v);
        } else if (threadIdx.x == 9 && v != 0.0) atomicAdd(&st->part[14], v);
    }
}

// one thread: prior, Jacobi scale, damping, gradient check, 3x3 Cholesky, candidate, model cost change
__global__ void icp_solve_kernel(IcpState* st) {
    LmState& s = st->lm;
    if (s.done) return;
    double* H = st->acc;   // 0 1 2 / 3 4 5 / 6 7 8 row-major (lower filled), g 9..11, cost 12, count 13, cand cost 14
    if (s.need_linearize) {
        for (int k = 0; k < 14; ++k) H[k] = st->part[k];
        H[1] = H[3]; H[2] = H[6]; H[5] = H[7];
        if (st->prior_w >= 0.0) {
            // PoseErrorRPZ (pose_error.hpp:147-154, residual order roll pitch z) / PoseErrorYXY (:176-183):
            // both are weight * (x_k - target_k) up to a row permutation, so J^T J = w^2 I, J^T r = w^2 (x - t)
            const double w2 = st->prior_w * st->prior_w;
            for (int k = 0; k < 3; ++k) { const double e = st->x[k] - st->prior_target[k]; H[4 * k] += w2; H[9 + k] += w2 * e; H[12] += 0.5 * w2 * e * e; }
        }
        s.cost_acc = H[12];
        s.n_accepted = (int)(H[13] + 0.5);
        double gm = 0.0;
        for (int k = 0; k < 3; ++k) { st->grad[k] = H[9 + k]; const double g = fabs(H[9 + k]); gm = (g == g) ? fmax(gm, g) : INFINITY; }
        s.grad_max_bits = (unsigned long long)__double_as_longlong(gm);
        if (!s.scale_valid) for (int k = 0; k < 3; ++k) st->scale[k] = s.jacobi ? 1.0 / (1.0 + sqrt(H[4 * k])) : 1.0;
    }
    for (int k = 0; k < 15; ++k) st->part[k] = 0.0;
    lm_control_pre(s);
    if (s.done) return;
    double A[9];
    for (int k = 0; k < 9; ++k) A[k] = H[k];
    for (int k = 0; k < 3; ++k) {
        const double s2 = st->scale[k] * st->scale[k];
        st->lam[k] = fmin(fmax(s2 * H[4 * k], s.min_diag), s.max_diag) / (s.radius * s2);
        A[4 * k] += st->lam[k];
    }
    // Cholesky 3x3
    double L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool ok = true;
    for (int j = 0; j < 3 && ok; ++j) {
        double dgn = A[4 * j]; for (int k = 0; k < j; ++k) dgn -= L[3 * j + k] * L[3 * j + k];
        if (!(dgn > 0.0)) { ok = false; break; }
        L[4 * j] = sqrt(dgn);
        for (int i = j + 1; i < 3; ++i) { double v = A[3 * i + j]; for (int k = 0; k < j; ++k) v -= L[3 * i + k] * L[3 * j + k]; L[3 * i + j] = v / L[4 * j]; }
    }
    double y[3], dx[3] = {0, 0, 0};
    if (ok) {
        for (int i = 0; i < 3; ++i) { double v = -st->grad[i]; for (int k = 0; k < i; ++k) v -= L[3 * i + k] * y[k]; y[i] = v / L[4 * i]; }
        for (int i = 2; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < 3; ++k) v -= L[3 * k + i] * dx[k]; dx[i] = v / L[4 * i]; }
    } else s.solve_fail = 1;
    double xn = 0, sn = 0, ma = 0, mb = 0;
    for (int k = 0; k < 3; ++k) {
        st->delta[k] = dx[k]; st->cand[k] = st->x[k] + dx[k];
        xn += st->x[k] * st->x[k]; sn += dx[k] * dx[k]; ma += dx[k] * st->lam[k] * dx[k]; mb += st->grad[k] * dx[k];
    }
    s.x_norm2 = xn; s.step_norm2 = sn; s.mcc_a = ma; s.mcc_b = mb;
}

__global__ void icp_post_kernel(IcpState* st) {
    LmState& s = st->lm;
    if (s.done) return;
    double cand = st->part[14];
    st->part[14] = 0.0;
    if (st->prior_w >= 0.0) { const double w2 = st->prior_w * st->prior_w; for (int k = 0; k < 3; ++k) { const double e = st->cand[k] - st->prior_target[k]; cand += 0.5 * w2 * e * e; } }
    s.cand_cost_acc = cand;
    lm_control_post(s);
    if (s.accept) {
        s.x_cost = cand;
        for (int k = 0; k < 3; ++k) st->x[k] = st->cand[k];
    }
}

}  // namespace

// ====================================================================================== host side
struct lvb_icp {
    lvb_ctx* ctx = nullptr;
    bool have_map = false;
    int P = 0;
    float cell_size = 0;
    Grid grid;
    DevBuf<unsigned char> map_raw, scan_raw, accepted;
    DevBuf<float4> map_sorted;
    DevBuf<int> cell_of, counts, cell_start, fill, block_sums, bbox, total, coarse;
    int cbx = 0, cby = 0, cbz = 0;
    DevBuf<float> pa;
    DevBuf<double> nrm, eval_r, eval_J;
    DevBuf<int> knn_idx;
    DevBuf<float> knn_d2;
    DevBuf<IcpState> st;
    DevBuf<int> q_key, q_counts, q_start, q_fill, q_sums, q_total, q_order;
    // device-resident world clouds per keyframe (Mapping::pointclouds_surf / _ground, mapping.cpp:208-219) and the merged map frame
    struct Segment { long long key; DevBuf<float4>* pts; int n; };
    std::vector<Segment> segments;
    DevBuf<float4> merged, merged2;
    int scan_n = 0, scan_stride = 0;          // the scan of the last upload_scan (reusable by lvb_icp_map_append)
    ~lvb_icp() { for (auto& sg : segments) delete sg.pts; }
};

#define ILAUNCH(h, kernel, grid, block, ...)                                              \
    do { if ((grid) > 0) { kernel<<<(grid), (block), 0, (h)->ctx->stream>>>(__VA_ARGS__); (h)->ctx->launches++; lvb::timing_mark((h)->ctx->stream, #kernel); } } while (0)

static inline int inblk(size_t n, int per) { return (int)((n + per - 1) / per); }
static int icheck(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch failed in %s: %s", what, cudaGetErrorString(e)); return LVB_ERR_CUDA; }
    return LVB_OK;
}

static int upload_scan(lvb_icp* h, const void* scan, int n, int stride, const double* frame_pose, float max_d2, double thr, IcpDev& d) {
    if (!h->have_map) { set_error("set_map first"); return LVB_ERR_STATE; }
    if (n < 0 || stride < 12 || (stride & 3) || (n && !scan)) { set_error("bad scan arguments"); return LVB_ERR_INVALID; }
    cudaStream_t s = h->ctx->stream;
    LVB_TRY(h->scan_raw.upload((const unsigned char*)scan, (size_t)n * stride, s));
    h->scan_n = n; h->scan_stride = stride;
    LVB_TRY(h->accepted.ensure(std::max(1, n)));
    LVB_TRY(h->pa.ensure((size_t)std::max(1, n) * 3));
    LVB_TRY(h->nrm.ensure((size_t)std::max(1, n) * 3));
    LVB_TRY(h->st.ensure(1));
    d.map = h->map_sorted.p; d.cell_start = h->cell_start.p; d.g = h->grid; d.P = h->P;
    d.coarse = h->coarse.p; d.cbx = h->cbx; d.cby = h->cby; d.cbz = h->cbz;
    d.scan = h->scan_raw.p; d.K = n; d.stride = stride;
    for (int i = 0; i < 7; ++i) d.tf[i] = (float)frame_pose[i];
    d.max_d2 = max_d2; d.thr = thr;
    d.accepted = h->accepted.p; d.pa = h->pa.p; d.nrm = h->nrm.p; d.st = h->st.p;
    d.rank = h->ctx->rank; d.world = h->ctx->world;
    d.order = nullptr;
    if (n >= 4096) {       // spatial visiting order: counting sort of the queries by coarse voxel block
        const Grid& g = h->grid;
        // key granularity: blocks of 2 x 2 x 2 voxels (measured on configs[2], association kernel: 4 x 4 x 4 blocks 0.51 ms, 2 x 2 x 2
        // 0.44 ms, single voxels 0.50 ms plus a slower sort) -- coarser only when the grid would need more than 8 M counters;
        // env LVB_ICP_SORT_SHIFT forces one (A/B)
        static const int forced = getenv("LVB_ICP_SORT_SHIFT") ? atoi(getenv("LVB_ICP_SORT_SHIFT")) : -1;
        int shift = 1;
        while (forced < 0 && (double)((g.gx >> shift) + 1) * ((g.gy >> shift) + 1) * ((g.gz >> shift) + 1) > 8.0e6) ++shift;
        if (forced >= 0) shift = forced;
        const int rnd = (1 << shift) - 1;
        const int cgx = (g.gx + rnd) >> shift, cgy = (g.gy + rnd) >> shift, cgz = (g.gz + rnd) >> shift;
        const int nc = cgx * cgy * cgz, nb = (nc + 1023) / 1024;
        LVB_TRY(h->q_key.ensure(n)); LVB_TRY(h->q_order.ensure(n)); LVB_TRY(h->q_counts.ensure(nc)); LVB_TRY(h->q_fill.ensure(nc));
        LVB_TRY(h->q_start.ensure((size_t)nc + 1)); LVB_TRY(h->q_sums.ensure(nb)); LVB_TRY(h->q_total.ensure(1));
        LVB_CUDA(cudaMemsetAsync(h->q_counts.p, 0, (size_t)nc * sizeof(int), s));
        LVB_CUDA(cudaMemsetAsync(h->q_fill.p, 0, (size_t)nc * sizeof(int), s));
        ILAUNCH(h, icp_query_key_kernel, (n + 255) / 256, 256, d, shift, cgx, cgy, cgz, h->q_key.p, h->q_counts.p);
        ILAUNCH(h, scan_block_kernel, nb, 1024, h->q_counts.p, h->q_start.p, nc, h->q_sums.p);
        ILAUNCH(h, scan_sums_kernel, 1, 1024, h->q_sums.p, nb, h->q_total.p);
        ILAUNCH(h, scan_add_kernel, nb, 1024, h->q_start.p, nc, h->q_sums.p, h->q_start.p + nc, h->q_total.p);
        ILAUNCH(h, icp_query_scatter_kernel, (n + 255) / 256, 256, n, h->q_key.p, h->q_start.p, h->q_fill.p, h->q_order.p);
        d.order = h->q_order.p;
    }
    return LVB_OK;
}

extern "C" {

int lvb_icp_create(lvb_ctx* ctx, lvb_icp** out) {
    if (!ctx || !out) { set_error("null argument"); return LVB_ERR_INVALID; }
    lvb_icp* h = new lvb_icp();
    h->ctx = ctx;
    *out = h;
    return LVB_OK;
}
void lvb_icp_destroy(lvb_icp* icp) { if (icp) { cudaSetDevice(icp->ctx->device); delete icp; } }

// voxel hash over the n records (stride bytes each) at d_points (device): bounding box, grid, counting sort by voxel key
static int build_hash(lvb_icp* h, const unsigned char* d_points, int n, int stride, float cell_size) {
    lvb_ctx* ctx = h->ctx;
    cudaStream_t s = ctx->stream;
    LVB_TRY(h->bbox.ensure(6));
    int init[6];
    { float big = FLT_MAX, small = -FLT_MAX; int bi, si; memcpy(&bi, &big, 4); memcpy(&si, &small, 4);
      const int so = si >= 0 ? si : si ^ 0x7fffffff; init[0] = init[1] = init[2] = bi; init[3] = init[4] = init[5] = so; }
    LVB_CUDA(cudaMemcpyAsync(h->bbox.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
    ILAUNCH(h, icp_bbox_kernel, std::min(1024, inblk(n, 256)), 256, d_points, n, stride, h->bbox.p);
    LVB_TRY(icheck("bbox"));
    int hb[6];
    LVB_TRY(h->bbox.download(hb, 6, s));
    LVB_CUDA(cudaStreamSynchronize(s));
    Grid g;
    // The voxel edge is (radius * 1.0001) / ring: the hair of slack guarantees that float rounding of the voxel
    // coordinate never separates two points closer than the radius by more than `ring` voxels.  ring = 4 unless the
    // bounding box would need more than MAX_CELLS voxels.
    g.minx = ord2f(hb[0]); g.miny = ord2f(hb[1]); g.minz = ord2f(hb[2]);
    const float mx = ord2f(hb[3]), my = ord2f(hb[4]), mz = ord2f(hb[5]);
    if (!(mx >= g.minx) || !(my >= g.miny) || !(mz >= g.minz) || !std::isfinite(mx - g.minx) || !std::isfinite(my - g.miny) || !std::isfinite(mz - g.minz)) {
        set_error("map cloud has no finite bounding box"); return LVB_ERR_INVALID; }
    double dx = 0, dy = 0, dz = 0;
    int ring = 4;
    for (;; ring >>= 1) {
        g.cell = cell_size * 1.0001f / (float)ring; g.inv_cell = 1.0f / g.cell;
        dx = std::floor((double)(mx - g.minx) * g.inv_cell) + 1; dy = std::floor((double)(my - g.miny) * g.inv_cell) + 1; dz = std::floor((double)(mz - g.minz) * g.inv_cell) + 1;
        if (dx * dy * dz <= (double)MAX_CELLS || ring == 1) break;
    }
    if (dx * dy * dz > (double)MAX_CELLS) { set_error("voxel grid of %.0f cells exceeds %d; increase cell_size or crop the map", dx * dy * dz, (int)MAX_CELLS); return LVB_ERR_UNSUPPORTED; }
    g.ring = ring;
    g.gx = (int)dx; g.gy = (int)dy; g.gz = (int)dz;
    const int ncell = g.gx * g.gy * g.gz;
    LVB_TRY(h->cell_of.ensure(n)); LVB_TRY(h->counts.ensure(ncell)); LVB_TRY(h->fill.ensure(ncell));
    LVB_TRY(h->cell_start.ensure((size_t)ncell + 1)); LVB_TRY(h->map_sorted.ensure(n));
    const int nb = inblk(ncell, 1024);
    LVB_TRY(h->block_sums.ensure(nb)); LVB_TRY(h->total.ensure(1));
    LVB_CUDA(cudaMemsetAsync(h->counts.p, 0, (size_t)ncell * sizeof(int), s));
    LVB_CUDA(cudaMemsetAsync(h->fill.p, 0, (size_t)ncell * sizeof(int), s));
    h->cbx = (g.gx + 3) >> 2; h->cby = (g.gy + 3) >> 2; h->cbz = (g.gz + 3) >> 2;
    LVB_TRY(h->coarse.ensure((size_t)h->cbx * h->cby * h->cbz));
    LVB_CUDA(cudaMemsetAsync(h->coarse.p, 0, (size_t)h->cbx * h->cby * h->cbz * sizeof(int), s));
    ILAUNCH(h, icp_count_kernel, inblk(n, 256), 256, d_points, n, stride, g, h->cell_of.p, h->counts.p, h->coarse.p, h->cbx, h->cby);
    ILAUNCH(h, scan_block_kernel, nb, 1024, h->counts.p, h->cell_start.p, ncell, h->block_sums.p);
    ILAUNCH(h, scan_sums_kernel, 1, 1024, h->block_sums.p, nb, h->total.p);
    ILAUNCH(h, scan_add_kernel, nb, 1024, h->cell_start.p, ncell, h->block_sums.p, h->cell_start.p + ncell, h->total.p);
    ILAUNCH(h, icp_scatter_kernel, inblk(n, 256), 256, d_points, n, stride, h->cell_of.p, h->cell_start.p, h->fill.p, h->map_sorted.p);
    LVB_TRY(icheck("set_map"));
    LVB_CUDA(cudaStreamSynchronize(s));
    h->grid = g; h->P = n; h->cell_size = cell_size; h->have_map = true;
    return LVB_OK;
}

int lvb_icp_set_map(lvb_icp* h, const void* points, int n, int stride, float cell_size) {
    if (n <= 0 || !points || stride < 12 || (stride & 3) || !(cell_size > 0.0f)) { set_error("bad map arguments"); return LVB_ERR_INVALID; }
    lvb_ctx* ctx = h->ctx;
    LVB_CUDA(cudaSetDevice(ctx->device)); lvb::g_alloc_stream = ctx->stream;
    LVB_TRY(h->map_raw.upload((const unsigned char*)points, (size_t)n * stride, ctx->stream));
    return build_hash(h, h->map_raw.p, n, stride, cell_size);
}

// ---- device-resident map: the per-keyframe world clouds stay in HBM, the map frame is merged and hashed there
int lvb_icp_map_append(lvb_icp* h, long long key, const void* robot_points, int n, int stride, const double pose[7]) {
    if (!h || !pose) { set_error("null argument"); return LVB_ERR_INVALID; }
    lvb_ctx* ctx = h->ctx;
    LVB_CUDA(cudaSetDevice(ctx->device)); lvb::g_alloc_stream = ctx->stream;
    cudaStream_t s = ctx->stream;
    const unsigned char* src = nullptr;
    DevBuf<unsigned char> up;
    if (robot_points) {
        if (n < 0 || stride < 12 || (stride & 3)) { set_error("bad cloud arguments"); return LVB_ERR_INVALID; }
        LVB_TRY(up.upload((const unsigned char*)robot_points, (size_t)n * stride, s));
        src = up.p;
    } else {      // the scan the last scan_to_map / knn3 / eval call uploaded: Mapping::Optimize followed by ToWorld of the same frame
        if (h->scan_n <= 0) { set_error("lvb_icp_map_append: no scan on the device"); return LVB_ERR_STATE; }
        n = h->scan_n; stride = h->scan_stride; src = h->scan_raw.p;
    }
    for (auto& sg : h->segments) if (sg.key == key) { set_error("lvb_icp_map_append: key %lld exists", key); return LVB_ERR_INVALID; }
    lvb_icp::Segment sg; sg.key = key; sg.n = n; sg.pts = new DevBuf<float4>();
    int rc = sg.pts->ensure(std::max(1, n));
    DevBuf<float> d_tf;
    float tf[7]; for (int i = 0; i < 7; ++i) tf[i] = (float)pose[i];      // Twc.cast<float>() (mapping.cpp:195)
    if (rc == LVB_OK) rc = d_tf.upload(tf, 7, s);
    if (rc != LVB_OK) { delete sg.pts; return rc; }
    ILAUNCH(h, icp_to_world_kernel, inblk(n, 256), 256, src, sg.pts->p, n, stride, d_tf.p);
    rc = icheck("map_append");
    if (rc != LVB_OK) { delete sg.pts; return rc; }
    LVB_CUDA(cudaStreamSynchronize(s));
    h->segments.push_back(sg);
    return LVB_OK;
}

int lvb_icp_map_evict(lvb_icp* h, long long key) {
    if (!h) { set_error("null argument"); return LVB_ERR_INVALID; }
    LVB_CUDA(cudaSetDevice(h->ctx->device)); lvb::g_alloc_stream = h->ctx->stream;
    size_t w = 0;
    for (size_t i = 0; i < h->segments.size(); ++i) {
        if (key < 0 || h->segments[i].key == key) delete h->segments[i].pts; else h->segments[w++] = h->segments[i];
    }
    h->segments.resize(w);
    return LVB_OK;
}

int lvb_icp_map_build(lvb_icp* h, const long long* keys, int n_keys, float cell_size, double ground_threshold, int* n_points) {
    if (!h || n_keys <= 0 || !keys || !(cell_size > 0.0f)) { set_error("bad arguments"); return LVB_ERR_INVALID; }
    lvb_ctx* ctx = h->ctx;
    LVB_CUDA(cudaSetDevice(ctx->device)); lvb::g_alloc_stream = ctx->stream;
    cudaStream_t s = ctx->stream;
    size_t total = 0;
    std::vector<const lvb_icp::Segment*> pick;
    for (int k = 0; k < n_keys; ++k) {
        const lvb_icp::Segment* f = nullptr;
        for (auto& sg : h->segments) if (sg.key == keys[k]) f = &sg;
        if (!f) { set_error("lvb_icp_map_build: key %lld is not resident", keys[k]); return LVB_ERR_INVALID; }
        pick.push_back(f); total += f->n;
    }
    if (total == 0 || total > 0x7fffffffull) { set_error("lvb_icp_map_build: %zu points", total); return LVB_ERR_INVALID; }
    // points_*_merged += pointclouds_*[time] in key order (mapping.cpp:121-125): the position in the merged cloud is the point's index
    LVB_TRY(h->merged.ensure(total));
    size_t off = 0;
    for (auto* f : pick) { if (f->n) LVB_CUDA(cudaMemcpyAsync(h->merged.p + off, f->pts->p, (size_t)f->n * sizeof(float4), cudaMemcpyDeviceToDevice, s)); off += f->n; }
    const float4* cloud = h->merged.p;
    int n = (int)total;
    if (ground_threshold > 0.0) {      // association_->SegmentGround(points_ground_merged) (mapping.cpp:126)
        LVB_TRY(h->merged2.ensure(total));
        LVB_TRY(lidar_segment_ground_device(ctx, h->merged.p, n, ground_threshold, h->merged2.p, &n));
        lvb::g_alloc_stream = ctx->stream;
        cloud = h->merged2.p;
        if (n <= 0) { set_error("lvb_icp_map_build: the ground plane fit kept no point"); return LVB_ERR_NUMERIC; }
    }
    if (n_points) *n_points = n;
    return build_hash(h, reinterpret_cast<const unsigned char*>(cloud), n, 16, cell_size);
}

int lvb_icp_map_download(lvb_icp* h, float* xyzi, int capacity, int* n_out) {
    if (!h || !h->have_map) { set_error("no map"); return LVB_ERR_STATE; }
    LVB_CUDA(cudaSetDevice(h->ctx->device));
    if (n_out) *n_out = h->P;
    if (!xyzi) return LVB_OK;
    if (capacity < h->P) { set_error("capacity %d < %d map points", capacity, h->P); return LVB_ERR_INVALID; }
    // hand the cloud back in its original (merge) order: map_sorted carries the original index in .w
    std::vector<float4> tmp(h->P);
    LVB_CUDA(cudaMemcpyAsync(tmp.data(), h->map_sorted.p, (size_t)h->P * sizeof(float4), cudaMemcpyDeviceToHost, h->ctx->stream));
    LVB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    for (int j = 0; j < h->P; ++j) { int i; memcpy(&i, &tmp[j].w, 4); if (i < 0 || i >= h->P) { set_error("corrupt index"); return LVB_ERR_NUMERIC; } xyzi[4 * (size_t)i] = tmp[j].x; xyzi[4 * (size_t)i + 1] = tmp[j].y; xyzi[4 * (size_t)i + 2] = tmp[j].z; xyzi[4 * (size_t)i + 3] = 0.0f; }
    return LVB_OK;
}

int lvb_icp_transform_cloud(lvb_icp* h, const void* points, int n, int stride, const double pose[7], void* out) {
    if (n < 0 || stride < 12 || (stride & 3) || (n && (!points || !out))) { set_error("bad cloud arguments"); return LVB_ERR_INVALID; }
    if (n == 0) return LVB_OK;
    LVB_CUDA(cudaSetDevice(h->ctx->device)); lvb::g_alloc_stream = h->ctx->stream;
    cudaStream_t s = h->ctx->stream;
    DevBuf<unsigned char> d_in, d_out; DevBuf<float> d_tf;
    float tf[7]; for (int i = 0; i < 7; ++i) tf[i] = (float)pose[i];      // Twc.cast<float>()
    LVB_TRY(d_in.upload((const unsigned char*)points, (size_t)n * stride, s)); LVB_TRY(d_out.ensure((size_t)n * stride)); LVB_TRY(d_tf.upload(tf, 7, s));
    ILAUNCH(h, icp_transform_cloud_kernel, inblk(n, 256), 256, d_in.p, d_out.p, n, stride, d_tf.p);
    LVB_TRY(icheck("transform_cloud"));
    LVB_TRY(d_out.download((unsigned char*)out, (size_t)n * stride, s));
    LVB_CUDA(cudaStreamSynchronize(s));
    return LVB_OK;
}

int lvb_icp_knn3(lvb_icp* h, const void* scan, int n, int stride, const double frame_pose[7], float max_d2, int32_t* idx, float* d2) {
    LVB_CUDA(cudaSetDevice(h->ctx->device)); lvb::g_alloc_stream = h->ctx->stream;
    if (h->have_map && max_d2 > h->cell_size * h->cell_size) { set_error("max_d2 %.6g exceeds cell_size^2 %.6g", max_d2, h->cell_size * h->cell_size); return LVB_ERR_INVALID; }
    IcpDev d;
    LVB_TRY(upload_scan(h, scan, n, stride, frame_pose, max_d2, 0.0, d));
    if (n == 0) return LVB_OK;
    LVB_TRY(h->knn_idx.ensure((size_t)n * 3)); LVB_TRY(h->knn_d2.ensure((size_t)n * 3));
    ILAUNCH(h, icp_knn_kernel, inblk(n, ITPB), ITPB, d, h->knn_idx.p, h->knn_d2.p);
    LVB_TRY(icheck("knn3"));
    LVB_TRY(h->knn_idx.download(idx, (size_t)n * 3, h->ctx->stream));
    LVB_TRY(h->knn_d2.download(d2, (size_t)n * 3, h->ctx->stream));
    LVB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return LVB_OK;
}

static int init_state(lvb_icp* h, int mode, const double* map_pose, const double* rpyxyz, double weight, double prior_w, double huber_a,
                      const lvb_solve_options* o) {
    lvb_solve_options opt;
    if (o) opt = *o; else lvb_default_options(&opt);
    IcpState hs;
    memset(&hs, 0, sizeof(hs));
    lm_init(hs.lm, opt);
    hs.mode = mode;
    memcpy(hs.Twc1, map_pose, sizeof(hs.Twc1)); memcpy(hs.rpyxyz, rpyxyz, sizeof(hs.rpyxyz));
    const int fr[2][3] = {{1, 2, 5}, {0, 3, 4}};
    for (int k = 0; k < 3; ++k) { hs.x[k] = rpyxyz[fr[mode][k]]; hs.cand[k] = hs.x[k]; hs.prior_target[k] = hs.x[k]; }
    hs.prior_w = prior_w; hs.huber_a = huber_a; hs.weight = weight;
    LVB_CUDA(cudaMemcpyAsync(h->st.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, h->ctx->stream));
    LVB_CUDA(cudaStreamSynchronize(h->ctx->stream));
    return LVB_OK;
}

int lvb_icp_eval(lvb_icp* h, int mode, const void* scan, int n, int stride, const double frame_pose[7], const double map_pose[7],
                 const double rpyxyz[6], double weight, double dist_thr, uint8_t* accepted, double* r, double* J) {
    if (mode < 0 || mode > 1) { set_error("bad mode"); return LVB_ERR_INVALID; }
    LVB_CUDA(cudaSetDevice(h->ctx->device)); lvb::g_alloc_stream = h->ctx->stream;
    IcpDev d;
    LVB_TRY(upload_scan(h, scan, n, stride, frame_pose, h->cell_size * h->cell_size, dist_thr, d));
    if (n == 0) return LVB_OK;
    LVB_TRY(init_state(h, mode, map_pose, rpyxyz, weight, -1.0, 0.0, nullptr));
    LVB_TRY(h->eval_r.ensure(n)); LVB_TRY(h->eval_J.ensure((size_t)n * 3));
    ILAUNCH(h, icp_associate_kernel, inblk(n, ITPB), ITPB, d);
    ILAUNCH(h, icp_eval_kernel, inblk(n, ITPB), ITPB, d, h->eval_r.p, h->eval_J.p);
    LVB_TRY(icheck("icp_eval"));
    cudaStream_t s = h->ctx->stream;
    if (accepted) LVB_TRY(h->accepted.download(accepted, n, s));
    if (r) LVB_TRY(h->eval_r.download(r, n, s));
    if (J) LVB_TRY(h->eval_J.download(J, (size_t)n * 3, s));
    LVB_CUDA(cudaStreamSynchronize(s));
    return LVB_OK;
}

int lvb_icp_scan_to_map(lvb_icp* h, int mode, const void* scan, int n, int stride, const double frame_pose[7], const double map_pose[7],
                        double rpyxyz[6], double weight, double prior_weight, double huber_a, double dist_thr,
                        const lvb_solve_options* options, lvb_solve_summary* summary) {
    if (mode < 0 || mode > 1) { set_error("bad mode"); return LVB_ERR_INVALID; }
    LVB_CUDA(cudaSetDevice(h->ctx->device)); lvb::g_alloc_stream = h->ctx->stream;
    const auto t0 = std::chrono::steady_clock::now();
    lvb_solve_options opt;
    if (options) opt = *options; else lvb_default_options(&opt);
    IcpDev d;
    lvb::timing_mark(h->ctx->stream, "begin");
    LVB_TRY(upload_scan(h, scan, n, stride, frame_pose, h->cell_size * h->cell_size, dist_thr, d));
    if ((double)h->cell_size * h->cell_size * (1.0 + 1e-6) < dist_thr) {
        set_error("dist_thr %.6g exceeds the map's cell_size^2 %.6g", dist_thr, (double)h->cell_size * h->cell_size); return LVB_ERR_INVALID; }
    LVB_TRY(init_state(h, mode, map_pose, rpyxyz, weight, prior_weight, huber_a, &opt));
    cudaStream_t s = h->ctx->stream;
    lvb_ctx* ctx = h->ctx;
    ILAUNCH(h, icp_associate_kernel, inblk(n, ITPB), ITPB, d);
    const int lin_blocks = std::max(1, std::min(inblk(n, ITPB), 4 * ctx->sm_count));
    IcpState hs;
    memset(&hs, 0, sizeof(hs));
    for (int pass = 0; pass <= opt.max_num_iterations; ++pass) {
        ILAUNCH(h, icp_linearize_kernel<0>, lin_blocks, ITPB, d);
        if (ctx->world > 1) LVB_TRY(comm_allreduce_sum_f64(ctx, d.st->part, 14));
        ILAUNCH(h, icp_solve_kernel, 1, 1, d.st);
        ILAUNCH(h, icp_linearize_kernel<1>, lin_blocks, ITPB, d);
        if (ctx->world > 1) LVB_TRY(comm_allreduce_sum_f64(ctx, d.st->part + 14, 1));
        ILAUNCH(h, icp_post_kernel, 1, 1, d.st);
        LVB_TRY(icheck("icp iteration"));
        LVB_CUDA(cudaMemcpyAsync(&hs, h->st.p, sizeof(hs), cudaMemcpyDeviceToHost, s));
        LVB_CUDA(cudaStreamSynchronize(s));
        LVB_TRY(comm_check(ctx));
        if (hs.lm.done) break;
        if (opt.max_solver_time_in_seconds < 1e8) {      // collective decision when the queries are sharded (see lvb_ba_solve)
            double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            LVB_TRY(comm_max_seconds(ctx, &el));
            if (el >= opt.max_solver_time_in_seconds) break;
        }
    }
    const int fr[2][3] = {{1, 2, 5}, {0, 3, 4}};
    for (int k = 0; k < 3; ++k) rpyxyz[fr[mode][k]] = hs.x[k];
    if (summary) {
        const int nb = hs.lm.n_accepted + (prior_weight >= 0 ? 1 : 0);
        summary->initial_cost = hs.lm.initial_cost; summary->final_cost = hs.lm.x_cost;
        summary->num_iterations = hs.lm.iter; summary->num_successful_steps = hs.lm.num_successful;
        summary->termination_type = hs.lm.termination; summary->num_residual_blocks = nb; summary->num_residual_blocks_reduced = nb;
        summary->final_radius = hs.lm.radius;
        summary->total_time_in_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return LVB_OK;
}

}  // extern "C"
