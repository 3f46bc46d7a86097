// oracle/ba.h -- sliding-window bundle-adjustment model of the CPU oracle
// (TEST INFRASTRUCTURE ONLY, parity unpinned).
//
// Holds the factor mix Backend::BuildProblem assembles
// (/root/reference/src/lvio_fusion/src/backend.cpp:96-183) in the same flat arrays the C ABI
// of the CUDA library takes, evaluates it with the functors of factors.h / imu.h, applies the
// loss + local parameterisation the way Ceres does [upstream], and solves the LM step by exact
// Schur elimination of the 1-dim inverse-depth blocks (what SPARSE_SCHUR does,
// backend.cpp:207) followed by a dense Cholesky of the reduced camera system.
#pragma once
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include "factors.h"
#include "imu.h"
#include "lm.h"

namespace oracle {

enum FactorKind { K_TWO_FRAME = 0, K_POSE_ONLY = 1, K_TWO_CAMERA = 2, K_IMU = 3, K_POSE_GRAPH = 4, K_POSE_PRIOR = 5, K_NUM = 6 };
static const int kConstStride[K_NUM] = {5, 6, 5, kImuConsts, 8, 9};
static const int kIdxStride[K_NUM] = {3, 1, 1, 8, 2, 1};
static const int kResDim[K_NUM] = {2, 2, 2, 15, 6, 6};
static const int kAmbientCols[K_NUM] = {15, 7, 1, 32, 14, 7};  // eval-mode Jacobian width

// [upstream] ceres::EigenQuaternionParameterization::Plus: q' = dq (x) q with
// dq = (sin|d|/|d| * d ; cos|d|), Eigen xyzw storage; translation/vec3/rho: plain addition.
inline void pose_plus(const double* x, const double* d, double* out) {
    const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (n > 0.0) {
        const double k = std::sin(n) / n;
        const Qd dq(k * d[0], k * d[1], k * d[2], std::cos(n));
        const Qd q = eig_mul(dq, Qd(x[0], x[1], x[2], x[3]));
        out[0] = q.x; out[1] = q.y; out[2] = q.z; out[3] = q.w;
    } else { for (int i = 0; i < 4; ++i) out[i] = x[i]; }
    for (int i = 0; i < 3; ++i) out[4 + i] = x[4 + i] + d[3 + i];
}
// [upstream] EigenQuaternionParameterization::ComputeJacobian, 4x3 row-major at q=(x,y,z,w)
inline void quat_plus_jacobian(const double* q, double* j) {
    j[0] = q[3];  j[1] = q[2];   j[2] = -q[1];
    j[3] = -q[2]; j[4] = q[3];   j[5] = q[0];
    j[6] = q[1];  j[7] = -q[0];  j[8] = q[3];
    j[9] = -q[0]; j[10] = -q[1]; j[11] = -q[2];
}
// ambient [rows x 7] (row stride ld) -> tangent [rows x 6]
inline void pose_to_tangent(const double* Ja, int rows, int ld, const double* pose, double* Jt /*rows x 6*/) {
    double pj[12];
    quat_plus_jacobian(pose, pj);
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < 3; ++c) {
            double s = 0; for (int k = 0; k < 4; ++k) s += Ja[r * ld + k] * pj[k * 3 + c];
            Jt[r * 6 + c] = s;
        }
        for (int c = 0; c < 3; ++c) Jt[r * 6 + 3 + c] = Ja[r * ld + 4 + c];
    }
}

// Persistent worker pool (Ceres keeps one too): parallel_for(T, fn) runs fn(0..T-1), task 0 on the caller.
class WorkerPool {
public:
    static WorkerPool& get() { static WorkerPool p; return p; }
    void run(int T, const std::function<void(int)>& fn) {
        if (T <= 1) { fn(0); return; }
        std::unique_lock<std::mutex> lk(m_);
        while ((int)workers_.size() < T - 1) { const int id = (int)workers_.size(); workers_.emplace_back([this, id] { loop(id); }); }
        fn_ = &fn; tasks_ = T; pending_ = T - 1; ++epoch_;
        cv_.notify_all();
        lk.unlock();
        fn(0);
        lk.lock();
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }
private:
    WorkerPool() {}
    ~WorkerPool() { { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++epoch_; } cv_.notify_all(); for (auto& w : workers_) w.join(); }
    void loop(int id) {
        unsigned long seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
            if (stop_) return;
            seen = epoch_;
            if (id + 1 < tasks_) { const auto* f = fn_; lk.unlock(); (*f)(id + 1); lk.lock(); if (--pending_ == 0) done_.notify_all(); }
        }
    }
    std::mutex m_; std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    const std::function<void(int)>* fn_ = nullptr;
    int tasks_ = 0, pending_ = 0; unsigned long epoch_ = 0; bool stop_ = false;
};
inline void parallel_for(int T, const std::function<void(int)>& fn) { WorkerPool::get().run(T, fn); }

struct FactorGroup {
    int n = 0;
    std::vector<double> consts;
    std::vector<int32_t> idx;
    double huber_a = 0.0;  // <= 0: no loss (NULL / TrivialLoss)
};

struct BaProblem {
    Camera cam[2];
    std::vector<double> poses, vec3, rho;            // 7n, 3n, n
    std::vector<uint8_t> pose_const, vec3_const, rho_const;
    FactorGroup grp[K_NUM];
    int num_threads = 1;

    // ---- structure (finalize) ----
    std::vector<int> pose_slot, vec3_slot, rho_slot;  // -1 when constant
    int n_pose_free = 0, n_vec3_free = 0, n_rho_free = 0, dimc = 0;
    std::vector<int> lm_start, lm_factor;             // CSR rho -> (kind<<28 | factor index)
    // ---- linearisation ----
    std::vector<double> Hcc, gc, Hll, gl;
    struct FacW { int slot1, slot2; double w1[6], w2[6], h, g; };
    std::vector<FacW> fw_tf;                          // per TwoFrame factor
    std::vector<double> h_tc, g_tc;                   // per TwoCamera factor
    // ---- candidate ----
    std::vector<double> c_poses, c_vec3, c_rho;
    // scratch for solve
    std::vector<double> dl;

    int n_poses() const { return (int)poses.size() / 7; }
    int n_vec3() const { return (int)vec3.size() / 3; }
    int n_rho() const { return (int)rho.size(); }
    int pose_off(int i) const { return pose_slot[i] < 0 ? -1 : 6 * pose_slot[i]; }
    int vec3_off(int i) const { return vec3_slot[i] < 0 ? -1 : 6 * n_pose_free + 3 * vec3_slot[i]; }

    void finalize() {
        auto slots = [](const std::vector<uint8_t>& c, int n, std::vector<int>& s) {
            s.assign(n, -1); int k = 0; for (int i = 0; i < n; ++i) if (c.empty() || !c[i]) s[i] = k++; return k; };
        n_pose_free = slots(pose_const, n_poses(), pose_slot);
        n_vec3_free = slots(vec3_const, n_vec3(), vec3_slot);
        n_rho_free = slots(rho_const, n_rho(), rho_slot);
        dimc = 6 * n_pose_free + 3 * n_vec3_free;
        // CSR landmark -> factors
        std::vector<int> cnt(n_rho() + 1, 0);
        const FactorGroup& tf = grp[K_TWO_FRAME]; const FactorGroup& tc = grp[K_TWO_CAMERA];
        for (int f = 0; f < tf.n; ++f) cnt[tf.idx[3 * f] + 1]++;
        for (int f = 0; f < tc.n; ++f) cnt[tc.idx[f] + 1]++;
        lm_start.assign(n_rho() + 1, 0);
        for (int i = 0; i < n_rho(); ++i) lm_start[i + 1] = lm_start[i] + cnt[i + 1];
        lm_factor.assign(lm_start.back(), 0);
        std::vector<int> fill(lm_start.begin(), lm_start.end() - 1);
        for (int f = 0; f < tf.n; ++f) lm_factor[fill[tf.idx[3 * f]]++] = f;
        for (int f = 0; f < tc.n; ++f) lm_factor[fill[tc.idx[f]]++] = (1 << 28) | f;
        c_poses = poses; c_vec3 = vec3; c_rho = rho;
    }
    int dim() const { return dimc + n_rho_free; }

    // ------------------------------------------------------------------ raw evaluation
    // Residual + ambient Jacobian of factor f of `kind` at state (P,V,R).  J may be null.
    bool eval_factor(int kind, int f, const double* P, const double* V, const double* R, double* r, double* J) const {
        const FactorGroup& g = grp[kind];
        const double* c = &g.consts[(size_t)f * kConstStride[kind]];
        const int32_t* ix = &g.idx[(size_t)f * kIdxStride[kind]];
        switch (kind) {
        case K_TWO_FRAME: two_frame_eval(c, cam[0], cam[1], R[ix[0]], P + 7 * ix[1], P + 7 * ix[2], r, J); return true;
        case K_POSE_ONLY: pose_only_eval(c, cam[0], P + 7 * ix[0], r, J); return true;
        case K_TWO_CAMERA: two_camera_eval(c, cam[0], cam[1], R[ix[0]], r, J); return true;
        case K_POSE_GRAPH: pose_graph_eval(c, P + 7 * ix[0], P + 7 * ix[1], r, J); return true;
        case K_POSE_PRIOR: pose_prior_eval(c, P + 7 * ix[0], r, J); return true;
        case K_IMU: {
            const Preint pre = load_preint(c);
            static const double zero3[3] = {0.0, 0.0, 0.0};     // ImuInitError: Baj = Bgj = 0 (imu_error.hpp:141-142)
            const double* prm[8] = {P + 7 * ix[0], V + 3 * ix[1], V + 3 * ix[2], V + 3 * ix[3],
                                    P + 7 * ix[4], V + 3 * ix[5], ix[6] < 0 ? zero3 : V + 3 * ix[6], ix[7] < 0 ? zero3 : V + 3 * ix[7]};
            if (!J) return imu_error_evaluate(pre, prm, r, nullptr);
            double jb[8][15 * 7];
            double* jp[8]; for (int k = 0; k < 8; ++k) jp[k] = jb[k];
            if (!imu_error_evaluate(pre, prm, r, jp)) return false;
            if (pre.is_init()) { std::memset(jb[6], 0, sizeof(jb[6])); std::memset(jb[7], 0, sizeof(jb[7])); }
            static const int w[8] = {7, 3, 3, 3, 7, 3, 3, 3};
            int off = 0;
            for (int k = 0; k < 8; ++k) { for (int i = 0; i < 15; ++i) for (int j = 0; j < w[k]; ++j) J[i * 32 + off + j] = jb[k][i * w[k] + j]; off += w[k]; }
            return true; }
        }
        return false;
    }

    // cost = 1/2 sum rho(|r|^2) over all blocks at (P,V,R); multi-threaded over factors
    double total_cost(const double* P, const double* V, const double* R) const {
        const int T = std::max(1, num_threads);
        std::vector<double> part(T, 0.0);
        auto work = [&](int t) {
            double acc = 0.0;
            for (int kind = 0; kind < K_NUM; ++kind) {
                const FactorGroup& g = grp[kind];
                const int lo = (int)((int64_t)g.n * t / T), hi = (int)((int64_t)g.n * (t + 1) / T);
                for (int f = lo; f < hi; ++f) {
                    double r[15];
                    eval_factor(kind, f, P, V, R, r, nullptr);
                    double s = 0; for (int k = 0; k < kResDim[kind]; ++k) s += r[k] * r[k];
                    double rho, sr; huber(g.huber_a, s, &rho, &sr);
                    acc += 0.5 * rho;
                }
            }
            part[t] = acc;
        };
        run_threads(T, work);
        double c = 0; for (double p : part) c += p; return c;
    }

    template <class F> static void run_threads(int T, F&& work) { parallel_for(T, std::function<void(int)>(work)); }

    // ------------------------------------------------------------------ linearisation
    // Builds Hcc (dense dimc x dimc, full symmetric), gc, per-factor landmark couplings, Hll, gl.
    double linearize(std::vector<double>& g_out, std::vector<double>& hdiag) {
        const int T = std::max(1, num_threads);
        const size_t nn = (size_t)dimc * dimc;
        std::vector<std::vector<double>> Hs(T), gs(T);
        std::vector<double> costs(T, 0.0);
        fw_tf.resize(grp[K_TWO_FRAME].n);
        h_tc.assign(grp[K_TWO_CAMERA].n, 0.0); g_tc.assign(grp[K_TWO_CAMERA].n, 0.0);
        const double* P = poses.data(); const double* V = vec3.data(); const double* R = rho.data();

        auto work = [&](int t) {
            std::vector<double>& H = Hs[t]; std::vector<double>& gv = gs[t];
            H.assign(nn, 0.0); gv.assign(dimc, 0.0);
            double acc = 0.0;
            // generic accumulate of blocks: offs[k] (<0: constant) widths[k], tangent Jt[k] (rows x width)
            auto accumulate = [&](int rows, int nb, const int* offs, const int* wid, double* const* Jt, const double* r) {
                for (int a = 0; a < nb; ++a) {
                    if (offs[a] < 0) continue;
                    for (int i = 0; i < wid[a]; ++i) {
                        double s = 0; for (int k = 0; k < rows; ++k) s += Jt[a][k * wid[a] + i] * r[k];
                        gv[offs[a] + i] += s;
                    }
                    for (int b = 0; b < nb; ++b) {
                        if (offs[b] < 0) continue;
                        for (int i = 0; i < wid[a]; ++i) for (int j = 0; j < wid[b]; ++j) {
                            double s = 0; for (int k = 0; k < rows; ++k) s += Jt[a][k * wid[a] + i] * Jt[b][k * wid[b] + j];
                            H[(size_t)(offs[a] + i) * dimc + offs[b] + j] += s;
                        }
                    }
                }
            };
            for (int kind = 0; kind < K_NUM; ++kind) {
                const FactorGroup& g = grp[kind];
                const int lo = (int)((int64_t)g.n * t / T), hi = (int)((int64_t)g.n * (t + 1) / T);
                for (int f = lo; f < hi; ++f) {
                    double r[15], J[15 * 32];
                    eval_factor(kind, f, P, V, R, r, J);
                    const int rows = kResDim[kind], ld = kAmbientCols[kind];
                    double s = 0; for (int k = 0; k < rows; ++k) s += r[k] * r[k];
                    double rho_v, sr; huber(g.huber_a, s, &rho_v, &sr);
                    acc += 0.5 * rho_v;
                    for (int k = 0; k < rows; ++k) r[k] *= sr;
                    for (int k = 0; k < rows * ld; ++k) J[k] *= sr;
                    const int32_t* ix = &g.idx[(size_t)f * kIdxStride[kind]];
                    if (kind == K_TWO_FRAME) {
                        double J1[12], J2[12];
                        pose_to_tangent(J + 1, 2, 15, P + 7 * ix[1], J1);
                        pose_to_tangent(J + 8, 2, 15, P + 7 * ix[2], J2);
                        int offs[2] = {pose_off(ix[1]), pose_off(ix[2])}; int wid[2] = {6, 6};
                        double* jt[2] = {J1, J2};
                        accumulate(2, 2, offs, wid, jt, r);
                        FacW& w = fw_tf[f];
                        w.slot1 = offs[0]; w.slot2 = offs[1];
                        const double jr0 = J[0], jr1 = J[15];
                        w.h = jr0 * jr0 + jr1 * jr1; w.g = jr0 * r[0] + jr1 * r[1];
                        for (int i = 0; i < 6; ++i) { w.w1[i] = jr0 * J1[i] + jr1 * J1[6 + i]; w.w2[i] = jr0 * J2[i] + jr1 * J2[6 + i]; }
                    } else if (kind == K_POSE_ONLY) {
                        double J1[12]; pose_to_tangent(J, 2, 7, P + 7 * ix[0], J1);
                        int offs[1] = {pose_off(ix[0])}; int wid[1] = {6}; double* jt[1] = {J1};
                        accumulate(2, 1, offs, wid, jt, r);
                    } else if (kind == K_TWO_CAMERA) {
                        h_tc[f] = J[0] * J[0] + J[1] * J[1]; g_tc[f] = J[0] * r[0] + J[1] * r[1];
                    } else if (kind == K_POSE_GRAPH) {
                        double J1[36], J2[36];
                        pose_to_tangent(J, 6, 14, P + 7 * ix[0], J1);
                        pose_to_tangent(J + 7, 6, 14, P + 7 * ix[1], J2);
                        int offs[2] = {pose_off(ix[0]), pose_off(ix[1])}; int wid[2] = {6, 6}; double* jt[2] = {J1, J2};
                        accumulate(6, 2, offs, wid, jt, r);
                    } else if (kind == K_POSE_PRIOR) {
                        double J1[36]; pose_to_tangent(J, 6, 7, P + 7 * ix[0], J1);
                        int offs[1] = {pose_off(ix[0])}; int wid[1] = {6}; double* jt[1] = {J1};
                        accumulate(6, 1, offs, wid, jt, r);
                    } else if (kind == K_IMU) {
                        double Jb[8][15 * 6];
                        int offs[8], wid[8]; double* jt[8];
                        static const int col0[8] = {0, 7, 10, 13, 16, 23, 26, 29};
                        for (int b = 0; b < 8; ++b) {
                            jt[b] = Jb[b];
                            if (b == 0 || b == 4) { pose_to_tangent(J + col0[b], 15, 32, P + 7 * ix[b], Jb[b]); offs[b] = pose_off(ix[b]); wid[b] = 6; }
                            else { for (int k = 0; k < 15; ++k) for (int c = 0; c < 3; ++c) Jb[b][k * 3 + c] = J[k * 32 + col0[b] + c]; offs[b] = ix[b] < 0 ? -1 : vec3_off(ix[b]); wid[b] = 3; }
                        }
                        accumulate(15, 8, offs, wid, jt, r);
                    }
                }
            }
            costs[t] = acc;
        };
        run_threads(T, work);
        Hcc.assign(nn, 0.0); gc.assign(dimc, 0.0);
        double cost = 0;
        for (int t = 0; t < T; ++t) {
            for (size_t i = 0; i < nn; ++i) Hcc[i] += Hs[t][i];
            for (int i = 0; i < dimc; ++i) gc[i] += gs[t][i];
            cost += costs[t];
        }
        // landmark diagonals
        Hll.assign(n_rho(), 0.0); gl.assign(n_rho(), 0.0);
        for (int l = 0; l < n_rho(); ++l) for (int e = lm_start[l]; e < lm_start[l + 1]; ++e) {
            const int f = lm_factor[e] & ((1 << 28) - 1);
            if (lm_factor[e] >> 28) { Hll[l] += h_tc[f]; gl[l] += g_tc[f]; } else { Hll[l] += fw_tf[f].h; gl[l] += fw_tf[f].g; }
        }
        g_out.assign(dim(), 0.0); hdiag.assign(dim(), 0.0);
        for (int i = 0; i < dimc; ++i) { g_out[i] = gc[i]; hdiag[i] = Hcc[(size_t)i * dimc + i]; }
        for (int l = 0; l < n_rho(); ++l) if (rho_slot[l] >= 0) { g_out[dimc + rho_slot[l]] = gl[l]; hdiag[dimc + rho_slot[l]] = Hll[l]; }
        return cost;
    }

    // Reduced camera system S (dimc x dimc) and rhs b so that S dc = b, for damping lambda.
    void reduced_system(const std::vector<double>& lambda, std::vector<double>& S, std::vector<double>& b) const {
        S = Hcc; b.assign(dimc, 0.0);
        for (int i = 0; i < dimc; ++i) { S[(size_t)i * dimc + i] += lambda[i]; b[i] = -gc[i]; }
        for (int l = 0; l < n_rho(); ++l) {
            if (rho_slot[l] < 0) continue;
            const double hl = Hll[l] + lambda[dimc + rho_slot[l]];
            if (!(hl > 0.0)) continue;
            const double hinv = 1.0 / hl;
            int slots[64]; double w[64][6]; int ns = 0;   // merged by pose slot
            for (int e = lm_start[l]; e < lm_start[l + 1]; ++e) {
                if (lm_factor[e] >> 28) continue;
                const FacW& fw = fw_tf[lm_factor[e]];
                const int so[2] = {fw.slot1, fw.slot2}; const double* ww[2] = {fw.w1, fw.w2};
                for (int k = 0; k < 2; ++k) {
                    if (so[k] < 0) continue;
                    int j = 0; for (; j < ns; ++j) if (slots[j] == so[k]) break;
                    if (j == ns) { slots[ns] = so[k]; for (int i = 0; i < 6; ++i) w[ns][i] = 0; ++ns; }
                    for (int i = 0; i < 6; ++i) w[j][i] += ww[k][i];
                }
            }
            for (int a = 0; a < ns; ++a) {
                for (int i = 0; i < 6; ++i) b[slots[a] + i] += w[a][i] * gl[l] * hinv;
                for (int c = 0; c < ns; ++c) for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j)
                    S[(size_t)(slots[a] + i) * dimc + slots[c] + j] -= w[a][i] * w[c][j] * hinv;
            }
        }
    }

    bool solve(const std::vector<double>& lambda, std::vector<double>& delta) {
        std::vector<double> S, b;
        reduced_system(lambda, S, b);
        if (dimc > 0 && !cholesky_solve(S, dimc, b)) return false;
        delta.assign(dim(), 0.0);
        for (int i = 0; i < dimc; ++i) delta[i] = b[i];
        for (int l = 0; l < n_rho(); ++l) {
            if (rho_slot[l] < 0) continue;
            const double hl = Hll[l] + lambda[dimc + rho_slot[l]];
            double s = gl[l];
            for (int e = lm_start[l]; e < lm_start[l + 1]; ++e) {
                if (lm_factor[e] >> 28) continue;
                const FacW& fw = fw_tf[lm_factor[e]];
                if (fw.slot1 >= 0) for (int i = 0; i < 6; ++i) s += fw.w1[i] * delta[fw.slot1 + i];
                if (fw.slot2 >= 0) for (int i = 0; i < 6; ++i) s += fw.w2[i] * delta[fw.slot2 + i];
            }
            delta[dimc + rho_slot[l]] = -s / hl;
        }
        return true;
    }

    void plus(const std::vector<double>& d, std::vector<double>& P, std::vector<double>& V, std::vector<double>& R) const {
        P = poses; V = vec3; R = rho;
        for (int i = 0; i < n_poses(); ++i) if (pose_slot[i] >= 0) pose_plus(&poses[7 * i], &d[6 * pose_slot[i]], &P[7 * i]);
        for (int i = 0; i < n_vec3(); ++i) if (vec3_slot[i] >= 0) for (int k = 0; k < 3; ++k) V[3 * i + k] += d[6 * n_pose_free + 3 * vec3_slot[i] + k];
        for (int l = 0; l < n_rho(); ++l) if (rho_slot[l] >= 0) R[l] += d[dimc + rho_slot[l]];
    }
    double candidate_cost(const std::vector<double>& d) { plus(d, c_poses, c_vec3, c_rho); return total_cost(c_poses.data(), c_vec3.data(), c_rho.data()); }
    void accept() { poses = c_poses; vec3 = c_vec3; rho = c_rho; }

    template <class Fn> void for_free(const std::vector<double>& P, const std::vector<double>& V, const std::vector<double>& R, Fn fn) const {
        for (int i = 0; i < n_poses(); ++i) if (pose_slot[i] >= 0) for (int k = 0; k < 7; ++k) fn(i * 16 + k, P[7 * i + k], 0);
        for (int i = 0; i < n_vec3(); ++i) if (vec3_slot[i] >= 0) for (int k = 0; k < 3; ++k) fn(0, V[3 * i + k], 1);
        for (int l = 0; l < n_rho(); ++l) if (rho_slot[l] >= 0) fn(0, R[l], 2);
    }
    double x_norm() const { double s = 0; for_free(poses, vec3, rho, [&](int, double v, int) { s += v * v; }); return std::sqrt(s); }
    double diff_norm(const std::vector<double>& P, const std::vector<double>& V, const std::vector<double>& R, bool inf) const {
        double s = 0;
        auto upd = [&](double d) { if (inf) s = std::fmax(s, std::fabs(d)); else s += d * d; };
        for (int i = 0; i < n_poses(); ++i) if (pose_slot[i] >= 0) for (int k = 0; k < 7; ++k) upd(poses[7 * i + k] - P[7 * i + k]);
        for (int i = 0; i < n_vec3(); ++i) if (vec3_slot[i] >= 0) for (int k = 0; k < 3; ++k) upd(vec3[3 * i + k] - V[3 * i + k]);
        for (int l = 0; l < n_rho(); ++l) if (rho_slot[l] >= 0) upd(rho[l] - R[l]);
        return inf ? s : std::sqrt(s);
    }
    double step_norm() const { return diff_norm(c_poses, c_vec3, c_rho, false); }
    double gradient_max_norm(const std::vector<double>& g) const {
        std::vector<double> ng(g.size()), P, V, R;
        for (size_t i = 0; i < g.size(); ++i) ng[i] = -g[i];
        plus(ng, P, V, R);
        return diff_norm(P, V, R, true);
    }
};

}  // namespace oracle
