// oracle/oracle_capi.cpp -- C exports of the CPU oracle (TEST INFRASTRUCTURE ONLY).
//
// Mirrors include/lvio_b200.h with the prefix orc_ so that the same Python driver code can run
// a problem through the oracle and through the CUDA library and compare.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
// "parity unpinned": the reference ships no golden vectors and cannot be built here
// (Eigen, Sophus, Ceres, PCL, glog, OpenCV are all absent).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include "../include/lvio_b200.h"
#include "ba.h"
#include "icp.h"
#include "lidar.h"

using namespace oracle;

namespace oracle { void icp_parallel(int T, const std::function<void(int)>& fn) { parallel_for(T, fn); } }

struct orc_ctx { int dummy; };
struct orc_ba { BaProblem p; bool finalized = false; };
struct orc_icp {
    std::vector<Pt> map; KdTree tree; float cell = 0; int num_threads = 1; bool brute = false;
};

static thread_local std::string g_err;
static int fail(int code, const char* msg) { g_err = msg; return code; }

static LmOptions to_lm(const lvb_solve_options* o) {
    LmOptions l;
    if (!o) return l;
    l.max_num_iterations = o->max_num_iterations;
    l.max_solver_time_in_seconds = o->max_solver_time_in_seconds;
    l.function_tolerance = o->function_tolerance;
    l.gradient_tolerance = o->gradient_tolerance;
    l.parameter_tolerance = o->parameter_tolerance;
    l.initial_trust_region_radius = o->initial_trust_region_radius;
    l.jacobi_scaling = o->jacobi_scaling;
    return l;
}
static void from_lm(const LmSummary& s, int nblocks, lvb_solve_summary* out) {
    if (!out) return;
    out->initial_cost = s.initial_cost; out->final_cost = s.final_cost;
    out->num_iterations = s.num_iterations; out->num_successful_steps = s.num_successful_steps;
    out->termination_type = s.termination; out->num_residual_blocks = nblocks;
    out->num_residual_blocks_reduced = nblocks; out->final_radius = s.final_radius;
    out->total_time_in_seconds = s.total_time_s;
}

extern "C" {

int orc_version(void) { return 100; }
void orc_default_options(lvb_solve_options* o) {
    o->max_num_iterations = 50; o->max_solver_time_in_seconds = 1e9; o->function_tolerance = 1e-6;
    o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8; o->initial_trust_region_radius = 1e4;
    o->jacobi_scaling = 1; o->linear_solver_type = 0; o->num_threads = 1; o->schur_mode = 0;
}
int orc_ctx_create(int, void*, orc_ctx** out) { *out = new orc_ctx(); return LVB_OK; }
void orc_ctx_destroy(orc_ctx* c) { delete c; }
const char* orc_last_error(void) { return g_err.c_str(); }
long long orc_launch_count(orc_ctx*) { return 0; }
int orc_ctx_synchronize(orc_ctx*) { return LVB_OK; }

int orc_ba_create(orc_ctx*, orc_ba** out) { *out = new orc_ba(); return LVB_OK; }
void orc_ba_destroy(orc_ba* b) { delete b; }
int orc_ba_set_cameras(orc_ba* b, const double* cam) { b->p.cam[0] = load_camera(cam); b->p.cam[1] = load_camera(cam + 11); return LVB_OK; }
int orc_ba_set_poses(orc_ba* b, int n, const double* v, const uint8_t* c) {
    b->p.poses.assign(v, v + (size_t)7 * n); if (c) b->p.pose_const.assign(c, c + n); else b->p.pose_const.assign(n, 0); return LVB_OK; }
int orc_ba_set_vec3(orc_ba* b, int n, const double* v, const uint8_t* c) {
    b->p.vec3.assign(v, v + (size_t)3 * n); if (c) b->p.vec3_const.assign(c, c + n); else b->p.vec3_const.assign(n, 0); return LVB_OK; }
int orc_ba_set_inv_depths(orc_ba* b, int n, const double* v, const uint8_t* c) {
    b->p.rho.assign(v, v + n); if (c) b->p.rho_const.assign(c, c + n); else b->p.rho_const.assign(n, 0); return LVB_OK; }
int orc_ba_add_factors(orc_ba* b, int kind, int n, const double* consts, const int32_t* idx) {
    if (kind < 0 || kind >= K_NUM) return fail(LVB_ERR_INVALID, "bad kind");
    FactorGroup& g = b->p.grp[kind];
    g.consts.insert(g.consts.end(), consts, consts + (size_t)n * kConstStride[kind]);
    g.idx.insert(g.idx.end(), idx, idx + (size_t)n * kIdxStride[kind]);
    g.n += n; b->finalized = false; return LVB_OK;
}
int orc_ba_set_loss(orc_ba* b, int kind, double a) { if (kind < 0 || kind >= K_NUM) return fail(LVB_ERR_INVALID, "bad kind"); b->p.grp[kind].huber_a = a; return LVB_OK; }
int orc_ba_finalize(orc_ba* b) {
    BaProblem& p = b->p;
    const int np = p.n_poses(), nv = p.n_vec3(), nr = p.n_rho();
    for (int k = 0; k < K_NUM; ++k) {
        const FactorGroup& g = p.grp[k];
        for (int f = 0; f < g.n; ++f) for (int j = 0; j < kIdxStride[k]; ++j) {
            const int v = g.idx[(size_t)f * kIdxStride[k] + j];
            int lim = np;
            if ((k == K_TWO_FRAME && j == 0) || k == K_TWO_CAMERA) lim = nr;
            if (k == K_IMU && j != 0 && j != 4) lim = nv;
            if (k == K_IMU && (j == 6 || j == 7) && v == -1 && g.consts[(size_t)f * kConstStride[k] + 467] >= 0.0) continue;   // ImuInitError
            if (v < 0 || v >= lim) return fail(LVB_ERR_INVALID, "factor index out of range");
        }
    }
    p.finalize(); b->finalized = true; return LVB_OK;
}
int orc_ba_dims(orc_ba* b, int* dimc, int* nrf, int* rows) {
    if (!b->finalized) return fail(LVB_ERR_STATE, "finalize first");
    if (dimc) *dimc = b->p.dimc; if (nrf) *nrf = b->p.n_rho_free;
    if (rows) { int r = 0; for (int k = 0; k < K_NUM; ++k) r += b->p.grp[k].n * kResDim[k]; *rows = r; }
    return LVB_OK;
}
int orc_ba_update_params(orc_ba* b, const double* P, const double* V, const double* R) {
    if (P) b->p.poses.assign(P, P + b->p.poses.size());
    if (V) b->p.vec3.assign(V, V + b->p.vec3.size());
    if (R) b->p.rho.assign(R, R + b->p.rho.size());
    return LVB_OK;
}
int orc_ba_eval(orc_ba* b, int kind, double* r, double* J) {
    if (!b->finalized) return fail(LVB_ERR_STATE, "finalize first");
    const BaProblem& p = b->p; const FactorGroup& g = p.grp[kind];
    const int T = std::max(1, p.num_threads);
    bool ok = true;
    auto work = [&](int t) {
        const int lo = (int)((int64_t)g.n * t / T), hi = (int)((int64_t)g.n * (t + 1) / T);
        double rr[15], JJ[15 * 32];
        for (int f = lo; f < hi; ++f) {
            if (!p.eval_factor(kind, f, p.poses.data(), p.vec3.data(), p.rho.data(), rr, J ? JJ : nullptr)) ok = false;
            if (r) std::memcpy(r + (size_t)f * kResDim[kind], rr, sizeof(double) * kResDim[kind]);
            if (J) std::memcpy(J + (size_t)f * kResDim[kind] * kAmbientCols[kind], JJ, sizeof(double) * kResDim[kind] * kAmbientCols[kind]);
        }
    };
    BaProblem::run_threads(T, work);
    return ok ? LVB_OK : fail(LVB_ERR_NUMERIC, "factor evaluation failed");
}
int orc_ba_eval_device(orc_ba*, int) { return fail(LVB_ERR_UNSUPPORTED, "oracle has no device"); }
int orc_ba_set_threads(orc_ba* b, int t) { b->p.num_threads = std::max(1, t); return LVB_OK; }

int orc_ba_reduced_system(orc_ba* b, double radius, double* S, double* rhs, double* cost) {
    if (!b->finalized) return fail(LVB_ERR_STATE, "finalize first");
    BaProblem& p = b->p;
    std::vector<double> g, hd;
    const double c = p.linearize(g, hd);
    const int n = p.dim();
    std::vector<double> lambda(n);
    LmOptions o;
    for (int j = 0; j < n; ++j) {
        const double s = 1.0 / (1.0 + std::sqrt(hd[j])); const double s2 = s * s;
        const double d = std::fmin(std::fmax(s2 * hd[j], o.min_lm_diagonal), o.max_lm_diagonal);
        lambda[j] = d / (radius * s2);
    }
    std::vector<double> Sv, bv;
    p.reduced_system(lambda, Sv, bv);
    if (S) std::memcpy(S, Sv.data(), sizeof(double) * Sv.size());
    if (rhs) std::memcpy(rhs, bv.data(), sizeof(double) * bv.size());
    if (cost) *cost = c;
    return LVB_OK;
}
int orc_ba_solve(orc_ba* b, const lvb_solve_options* o, lvb_solve_summary* s) {
    if (!b->finalized) return fail(LVB_ERR_STATE, "finalize first");
    if (o && o->num_threads > 0) b->p.num_threads = o->num_threads;
    LmSummary sum; lm_minimize(b->p, to_lm(o), sum);
    int nb = 0; for (int k = 0; k < K_NUM; ++k) nb += b->p.grp[k].n;
    from_lm(sum, nb, s); return LVB_OK;
}
// one Gauss-Newton/LM pass with a fixed radius (linearize + solve + candidate cost + accept):
// the unit of work bench.py times as a "step" on the CPU side.
int orc_ba_iterate(orc_ba* b, int iters, double radius, double* cost_out) {
    BaProblem& p = b->p; std::vector<double> g, hd, lambda, delta; double c = 0;
    for (int it = 0; it < iters; ++it) {
        c = p.linearize(g, hd);
        const int n = p.dim(); lambda.resize(n);
        for (int j = 0; j < n; ++j) { const double s = 1.0 / (1.0 + std::sqrt(hd[j])); const double s2 = s * s;
            lambda[j] = std::fmin(std::fmax(s2 * hd[j], 1e-6), 1e32) / (radius * s2); }
        if (!p.solve(lambda, delta)) return fail(LVB_ERR_NUMERIC, "cholesky failed");
        const double cc = p.candidate_cost(delta);
        if (cc < c) p.accept();
    }
    if (cost_out) *cost_out = c;
    return LVB_OK;
}
int orc_ba_get_poses(orc_ba* b, double* o) { std::memcpy(o, b->p.poses.data(), sizeof(double) * b->p.poses.size()); return LVB_OK; }
int orc_ba_get_vec3(orc_ba* b, double* o) { std::memcpy(o, b->p.vec3.data(), sizeof(double) * b->p.vec3.size()); return LVB_OK; }
int orc_ba_get_inv_depths(orc_ba* b, double* o) { std::memcpy(o, b->p.rho.data(), sizeof(double) * b->p.rho.size()); return LVB_OK; }
int orc_ba_reprojection_errors(orc_ba* b, int n, const double* ob_pw, const int32_t* pose_idx, double* err) {
    for (int i = 0; i < n; ++i) {
        const double* e = ob_pw + 5 * i;
        const double c[6] = {e[0], e[1], e[2], e[3], e[4], 1.0};
        double r[2]; pose_only_eval(c, b->p.cam[0], &b->p.poses[7 * pose_idx[i]], r, nullptr);
        err[i] = std::sqrt(r[0] * r[0] + r[1] * r[1]);
    }
    return LVB_OK;
}

// ---- ICP ------------------------------------------------------------------------------
int orc_icp_create(orc_ctx*, orc_icp** out) { *out = new orc_icp(); return LVB_OK; }
void orc_icp_destroy(orc_icp* h) { delete h; }
int orc_icp_set_threads(orc_icp* h, int t) { h->num_threads = std::max(1, t); return LVB_OK; }
int orc_icp_set_brute(orc_icp* h, int b) { h->brute = b != 0; return LVB_OK; }
static Pt load_pt(const void* base, int i, int stride) { const float* f = (const float*)((const char*)base + (size_t)i * stride); return Pt{f[0], f[1], f[2]}; }
int orc_icp_set_map(orc_icp* h, const void* pts, int n, int stride, float cell) {
    if (stride < 12) return fail(LVB_ERR_INVALID, "stride < 12");
    h->map.resize(n); for (int i = 0; i < n; ++i) h->map[i] = load_pt(pts, i, stride);
    h->cell = cell; h->tree.build(h->map.data(), n); return LVB_OK;
}
static Knn3 query_limited(const orc_icp* h, const Pt& q, float max_d2) {
    Knn3 k = h->brute ? knn3_brute(h->map.data(), (int)h->map.size(), q) : h->tree.query(q);
    for (int j = 0; j < 3; ++j) if (!(k.d2[j] <= max_d2)) { k.idx[j] = -1; k.d2[j] = std::numeric_limits<float>::infinity(); }
    return k;
}
// Mapping::MergeScan (mapping.cpp:193-203)
int orc_icp_transform_cloud(orc_icp*, const void* pts, int n, int stride, const double* pose, void* out) {
    std::memcpy(out, pts, (size_t)n * stride);
    for (int i = 0; i < n; ++i) { const Pt q = transform_f32(pose, load_pt(pts, i, stride)); float* f = (float*)((char*)out + (size_t)i * stride); f[0] = q.x; f[1] = q.y; f[2] = q.z; }
    return LVB_OK;
}
int orc_icp_knn3(orc_icp* h, const void* scan, int n, int stride, const double* pose, float max_d2, int32_t* idx, float* d2) {
    const int T = std::max(1, h->num_threads);
    auto work = [&](int t) {
        const int lo = (int)((int64_t)n * t / T), hi = (int)((int64_t)n * (t + 1) / T);
        for (int i = lo; i < hi; ++i) {
            const Knn3 k = query_limited(h, transform_f32(pose, load_pt(scan, i, stride)), max_d2);
            for (int j = 0; j < 3; ++j) { idx[3 * i + j] = k.idx[j]; d2[3 * i + j] = k.d2[j]; }
        }
    };
    BaProblem::run_threads(T, work);
    return LVB_OK;
}
// association.cpp:291-317: gate + factor constants (p, pa, n, weight) for accepted points
static void associate(orc_icp* h, const void* scan, int n, int stride, const double* frame_pose, double weight, double thr,
                      std::vector<uint8_t>& acc, std::vector<double>& consts) {
    const int P = (int)h->map.size();
    const float max_d2 = h->cell * h->cell;
    acc.assign(n, 0); consts.assign((size_t)n * 10, 0.0);
    const int T = std::max(1, h->num_threads);
    auto work = [&](int t) {
        const int lo = (int)((int64_t)n * t / T), hi = (int)((int64_t)n * (t + 1) / T);
        for (int i = lo; i < hi; ++i) {
            const Pt p = load_pt(scan, i, stride);
            const Knn3 k = query_limited(h, transform_f32(frame_pose, p), max_d2);
            bool ok = true;
            for (int j = 0; j < 3; ++j) if (!(k.idx[j] >= 0 && k.idx[j] < P && (double)k.d2[j] < thr)) ok = false;
            if (!ok) continue;
            acc[i] = 1;
            const Pt a = h->map[k.idx[0]], b = h->map[k.idx[1]], c = h->map[k.idx[2]];
            const V3d pa(a.x, a.y, a.z), pb(b.x, b.y, b.z), pc(c.x, c.y, c.z);
            const V3d nn = plane_normal(pa, pb, pc);
            double* o = &consts[(size_t)i * 10];
            o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = pa.x; o[4] = pa.y; o[5] = pa.z; o[6] = nn.x; o[7] = nn.y; o[8] = nn.z; o[9] = weight;
        }
    };
    BaProblem::run_threads(T, work);
}
int orc_icp_eval(orc_icp* h, int mode, const void* scan, int n, int stride, const double* frame_pose, const double* map_pose,
                 const double* rpyxyz, double weight, double thr, uint8_t* accepted, double* r, double* J) {
    std::vector<uint8_t> acc; std::vector<double> consts;
    associate(h, scan, n, stride, frame_pose, weight, thr, acc, consts);
    int f[3]; IcpProblem::free_index(mode, f);
    const double x[3] = {rpyxyz[f[0]], rpyxyz[f[1]], rpyxyz[f[2]]};
    for (int i = 0; i < n; ++i) {
        if (accepted) accepted[i] = acc[i];
        double rr = 0, JJ[3] = {0, 0, 0};
        if (acc[i]) lidar_plane_eval(&consts[(size_t)i * 10], mode, map_pose, rpyxyz, x, &rr, JJ);
        if (r) r[i] = rr;
        if (J) for (int k = 0; k < 3; ++k) J[3 * i + k] = JJ[k];
    }
    return LVB_OK;
}
int orc_icp_scan_to_map(orc_icp* h, int mode, const void* scan, int n, int stride, const double* frame_pose, const double* map_pose,
                        double* rpyxyz, double weight, double prior_weight, double huber_a, double thr,
                        const lvb_solve_options* o, lvb_solve_summary* s) {
    std::vector<uint8_t> acc; std::vector<double> consts;
    associate(h, scan, n, stride, frame_pose, weight, thr, acc, consts);
    IcpProblem p; p.mode = mode; p.huber_a = huber_a; p.prior_weight = prior_weight; p.num_threads = h->num_threads;
    for (int i = 0; i < n; ++i) if (acc[i]) { p.consts.insert(p.consts.end(), &consts[(size_t)i * 10], &consts[(size_t)i * 10] + 10); ++p.n; }
    std::memcpy(p.Twc1, map_pose, sizeof(p.Twc1)); std::memcpy(p.rpyxyz, rpyxyz, sizeof(p.rpyxyz));
    p.get_free(p.prior_target);  // PoseErrorRPZ/YXY ctor captures the entries (pose_error.hpp:138-144,167-172)
    LmSummary sum; lm_minimize(p, to_lm(o), sum);
    std::memcpy(rpyxyz, p.rpyxyz, sizeof(p.rpyxyz));
    from_lm(sum, p.n + (prior_weight >= 0 ? 1 : 0), s);
    return LVB_OK;
}

// ---- producers / helpers only the oracle offers ---------------------------------------
// Preintegrate n IMU samples (dt, acc[3], gyr[3] rows of 7) starting from (acc0, gyr0) with
// linearisation biases; noise4 = ACC_N GYR_N ACC_W GYR_W.  out = 467 doubles (LVB_IMU consts).
int orc_preintegrate(int n, const double* samples7, const double* acc0, const double* gyr0,
                     const double* ba, const double* bg, const double* noise4, double* out) {
    PreintState s; preint_reset(s, V3d(ba[0], ba[1], ba[2]), V3d(bg[0], bg[1], bg[2]));
    const V3d a0(acc0[0], acc0[1], acc0[2]), g0(gyr0[0], gyr0[1], gyr0[2]);
    for (int i = 0; i < n; ++i) {
        const double* r = samples7 + 7 * i;
        preint_append(s, r[0], V3d(r[1], r[2], r[3]), V3d(r[4], r[5], r[6]), a0, g0, noise4);
    }
    store_preint(s.p, out); return LVB_OK;
}
// batch form with the product's signature (lvb_imu_preintegrate): interval i owns samples first[i]..first[i+1]-1
int orc_imu_preintegrate(void* /*ctx*/, int n, const int32_t* first, const double* samples7, const double* acc0, const double* gyr0,
                         const double* ba, const double* bg, const double* noise4, double* out) {
    if (n < 0 || (n && (!first || !acc0 || !gyr0 || !ba || !bg || !noise4 || !out))) return fail(LVB_ERR_INVALID, "orc_imu_preintegrate: bad arguments");
    for (int i = 0; i < n; ++i) {
        if (first[i + 1] < first[i]) return fail(LVB_ERR_INVALID, "orc_imu_preintegrate: first[] must be non-decreasing");
        orc_preintegrate(first[i + 1] - first[i], samples7 + 7 * (size_t)first[i], acc0 + 3 * i, gyr0 + 3 * i, ba + 3 * i, bg + 3 * i, noise4, out + (size_t)kImuConsts * i);
    }
    return LVB_OK;
}
int orc_sqrt_information(const double* cov225, double* U225) {
    double c[15][15], U[15][15]; std::memcpy(c, cov225, sizeof(c));
    if (!sqrt_information(c, U)) return fail(LVB_ERR_NUMERIC, "covariance inverse not SPD");
    std::memcpy(U225, U, sizeof(U)); return LVB_OK;
}
int orc_pose_plus(const double* x7, const double* d6, double* out7) { pose_plus(x7, d6, out7); return LVB_OK; }
int orc_pose_tangent(const double* Ja, int rows, int ld, const double* pose, double* Jt) { pose_to_tangent(Ja, rows, ld, pose, Jt); return LVB_OK; }
int orc_se3_to_rpyxyz(const double* T7, double* out6) { to_rpyxyz(load_rigid<double>(T7), out6); return LVB_OK; }
int orc_rpyxyz_to_se3(const double* in6, double* T7) { store_rigid(from_rpyxyz(in6), T7); return LVB_OK; }
int orc_se3_compose(const double* a, const double* b, double* out) { store_rigid(compose(load_rigid<double>(a), load_rigid<double>(b)), out); return LVB_OK; }
int orc_se3_inverse(const double* a, double* out) { store_rigid(inverse(load_rigid<double>(a)), out); return LVB_OK; }
int orc_transform_f32(const double* pose, int n, const float* in3, float* out3) {
    for (int i = 0; i < n; ++i) { const Pt q = transform_f32(pose, Pt{in3[3 * i], in3[3 * i + 1], in3[3 * i + 2]}); out3[3 * i] = q.x; out3[3 * i + 1] = q.y; out3[3 * i + 2] = q.z; }
    return LVB_OK;
}


// ---- lidar feature pipeline (lidar.h), same signatures as lvb_lidar_*
static LidarConfig to_cfg(const lvb_lidar_config* c) {
    LidarConfig k;
    k.num_scans = c->num_scans; k.horizon_scan = c->horizon_scan; k.ang_res_y = c->ang_res_y; k.ang_bottom = c->ang_bottom;
    k.ground_rows = c->ground_rows; k.cycle_time = c->cycle_time; k.min_range = c->min_range; k.max_range = c->max_range; k.resolution = c->resolution;
    std::memcpy(k.extrinsic, c->extrinsic, sizeof(k.extrinsic));
    return k;
}
static int bad_cfg(const lvb_lidar_config* c) {
    return !c || c->num_scans <= 0 || c->horizon_scan <= 0 || c->num_scans > 1024 || c->horizon_scan > 16384 || !(c->ang_res_y > 0) || c->ground_rows < 0 || !(c->resolution > 0);
}
static std::vector<PointI> to_cloud(const float* xyzi, int n) {
    std::vector<PointI> v(n);
    for (int i = 0; i < n; ++i) v[i] = PointI{xyzi[4 * i], xyzi[4 * i + 1], xyzi[4 * i + 2], xyzi[4 * i + 3]};
    return v;
}
static void from_cloud(const std::vector<PointI>& v, float* out, int32_t* n_out) {
    if (out) for (size_t i = 0; i < v.size(); ++i) { out[4 * i] = v[i].x; out[4 * i + 1] = v[i].y; out[4 * i + 2] = v[i].z; out[4 * i + 3] = v[i].intensity; }
    if (n_out) *n_out = (int32_t)v.size();
}
void orc_lidar_default_config(lvb_lidar_config* c) {
    std::memset(c, 0, sizeof(*c));
    c->num_scans = 64; c->horizon_scan = 1800; c->ang_res_y = 0.427; c->ang_bottom = 24.9; c->ground_rows = 60;
    c->cycle_time = 0.1036; c->min_range = 5; c->max_range = 30; c->resolution = 0.2; c->extrinsic[3] = 1.0;
}
int orc_lidar_segment(void*, const lvb_lidar_config* cfg, const void* points, int n, int stride, float* seg_xyzi, float* seg_range, uint8_t* seg_ground,
                      int32_t* seg_col, float* seg_curv, int32_t* start_ring, int32_t* end_ring, float* orientation, int32_t* n_seg) {
    if (bad_cfg(cfg) || n < 0 || (n && !points) || stride < 12 || (stride & 3)) return fail(LVB_ERR_INVALID, "orc_lidar_segment: bad arguments");
    const LidarConfig c = to_cfg(cfg);
    Segmented s = lidar_project_and_segment(c, lidar_preprocess(c, (const unsigned char*)points, n, stride));
    lidar_adjust_distortion(c, s);
    lidar_smoothness(s);
    from_cloud(s.pts, seg_xyzi, n_seg);
    for (size_t i = 0; i < s.pts.size(); ++i) {
        if (seg_range) seg_range[i] = s.range[i];
        if (seg_ground) seg_ground[i] = s.ground[i];
        if (seg_col) seg_col[i] = s.col[i];
        if (seg_curv) seg_curv[i] = s.curvature[i];
    }
    for (int i = 0; i < c.num_scans; ++i) { if (start_ring) start_ring[i] = s.start_ring[i]; if (end_ring) end_ring[i] = s.end_ring[i]; }
    if (orientation) { orientation[0] = s.start_orientation; orientation[1] = s.end_orientation; orientation[2] = s.orientation_diff; }
    return LVB_OK;
}
int orc_lidar_voxel_grid(void*, const float* xyzi, int n, float leaf, float* out, int32_t* n_out) {
    if (n < 0 || (n && !xyzi) || !(leaf > 0.0f)) return fail(LVB_ERR_INVALID, "orc_lidar_voxel_grid: bad arguments");
    from_cloud(voxel_grid(to_cloud(xyzi, n), leaf), out, n_out); return LVB_OK;
}
int orc_lidar_radius_outlier_removal(void*, const float* xyzi, int n, double radius, int min_neighbors, float* out, int32_t* n_out) {
    if (n < 0 || (n && !xyzi) || !(radius > 0.0)) return fail(LVB_ERR_INVALID, "orc_lidar_radius_outlier_removal: bad arguments");
    from_cloud(radius_outlier_removal(to_cloud(xyzi, n), radius, min_neighbors), out, n_out); return LVB_OK;
}
int orc_lidar_segment_ground(void*, const float* xyzi, int n, double thr, float* out, int32_t* n_out) {
    if (n < 0 || (n && !xyzi) || !(thr > 0.0)) return fail(LVB_ERR_INVALID, "orc_lidar_segment_ground: bad arguments");
    from_cloud(segment_ground(to_cloud(xyzi, n), thr), out, n_out); return LVB_OK;
}
int orc_lidar_extract_features(void*, const lvb_lidar_config* cfg, const void* points, int n, int stride, float* ground, int32_t* n_ground, float* surf, int32_t* n_surf) {
    if (bad_cfg(cfg) || n < 0 || (n && !points) || stride < 12 || (stride & 3)) return fail(LVB_ERR_INVALID, "orc_lidar_extract_features: bad arguments");
    std::vector<PointI> g, sf;
    lidar_extract_features(to_cfg(cfg), (const unsigned char*)points, n, stride, g, sf);
    from_cloud(g, ground, n_ground); from_cloud(sf, surf, n_surf);
    return LVB_OK;
}

}  // extern "C"
