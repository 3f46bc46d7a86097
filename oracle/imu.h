// oracle/imu.h -- IMU preintegration + ImuError restatement (TEST INFRASTRUCTURE ONLY).
// Pinned against the reference's own src/preintegration.cpp and ceres/imu_error.hpp compiled in place against a mini-Eigen
// (oracle/ref_harness.cpp, oracle/ref_compat/mini_eigen.h, tests/golden/ref_factors.npz, tests/test_ref_golden.py):
// preintegration record 1e-12, whitened residual / Jacobians 1e-8.  ImuInitError (the FullBA variant) is not in the fixture.
//
//   producer : /root/reference/src/lvio_fusion/src/preintegration.cpp:11-127 (midpoint propagate)
//   residual : /root/reference/src/lvio_fusion/src/preintegration.cpp:144-165 (Evaluate)
//   factor   : /root/reference/src/lvio_fusion/include/lvio_fusion/ceres/imu_error.hpp:12-122
//   helpers  : /root/reference/src/lvio_fusion/include/lvio_fusion/utility.h:99-140
//
// State ordering of the 15-vector: O_T=0 O_R=3 O_V=6 O_BA=9 O_BG=12; pose Jacobian columns
// O_PR=0 (rotation, 3 wide), column 3 identically zero, O_PT=4 (translation)
// (preintegration.cpp:11-12).  Gravity g = (0,0,+9.81007) (preintegration.cpp:13).
#pragma once
#include <cstring>
#include <vector>
#include "geometry.h"

namespace oracle {

static const double kGravity[3] = {0.0, 0.0, 9.81007};

// Preintegrated measurement as the C ABI ships it (467 doubles):
//   delta_p[3] delta_q[4 xyzw] delta_v[3] lin_ba[3] lin_bg[3] sum_dt jacobian[225 rm] covariance[225 rm] prior_a prior_g
// prior_a, prior_g < 0: ImuError (imu_error.hpp:12-122).  >= 0: ImuInitError (imu_error.hpp:124-229) -- the j-side biases
// are the constant zero vector, blocks (9,9) and (12,12) of the inverse covariance are replaced by prior * I.
enum { kImuConsts = 17 + 225 + 225 + 2 };

struct Preint {
    V3d dp, dv, lin_ba, lin_bg;
    Qd dq;
    double sum_dt;
    double jac[15][15];
    double cov[15][15];
    double prior_a = -1.0, prior_g = -1.0;
    bool is_init() const { return prior_a >= 0.0 && prior_g >= 0.0; }
};
inline Preint load_preint(const double* c) {
    Preint p;
    p.dp = V3d(c[0], c[1], c[2]);
    p.dq = Qd(c[3], c[4], c[5], c[6]);
    p.dv = V3d(c[7], c[8], c[9]);
    p.lin_ba = V3d(c[10], c[11], c[12]);
    p.lin_bg = V3d(c[13], c[14], c[15]);
    p.sum_dt = c[16];
    std::memcpy(p.jac, c + 17, sizeof(p.jac));
    std::memcpy(p.cov, c + 17 + 225, sizeof(p.cov));
    p.prior_a = c[467]; p.prior_g = c[468];
    return p;
}
inline void store_preint(const Preint& p, double* c) {
    c[0] = p.dp.x; c[1] = p.dp.y; c[2] = p.dp.z;
    c[3] = p.dq.x; c[4] = p.dq.y; c[5] = p.dq.z; c[6] = p.dq.w;
    c[7] = p.dv.x; c[8] = p.dv.y; c[9] = p.dv.z;
    c[10] = p.lin_ba.x; c[11] = p.lin_ba.y; c[12] = p.lin_ba.z;
    c[13] = p.lin_bg.x; c[14] = p.lin_bg.y; c[15] = p.lin_bg.z;
    c[16] = p.sum_dt;
    std::memcpy(c + 17, p.jac, sizeof(p.jac));
    std::memcpy(c + 17 + 225, p.cov, sizeof(p.cov));
    c[467] = p.prior_a; c[468] = p.prior_g;
}

inline Mat3 block3(const double (*m)[15], int r, int c) {
    Mat3 b; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) b.m[i][j] = m[r + i][c + j]; return b;
}
inline void set_block3(double (*m)[15], int r, int c, const Mat3& b) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[r + i][c + j] = b.m[i][j];
}

// utility.h:99-112 q_delta: (w=1, xyz=theta/2), NOT normalised
inline Qd q_delta(const V3d& theta) { return Qd(theta.x / 2.0, theta.y / 2.0, theta.z / 2.0, 1.0); }

// utility.h:124-140 q_left / q_right as full 4x4 (row/col 0 = w)
struct Mat4 { double m[4][4]; };
inline Mat4 q_left(const Qd& q) {
    Mat4 a; const V3d v(q.x, q.y, q.z); const Mat3 s = skew(v);
    a.m[0][0] = q.w;
    for (int j = 0; j < 3; ++j) { a.m[0][1 + j] = -v[j]; a.m[1 + j][0] = v[j]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a.m[1 + i][1 + j] = (i == j ? q.w : 0.0) + s.m[i][j];
    return a;
}
inline Mat4 q_right(const Qd& p) {
    Mat4 a; const V3d v(p.x, p.y, p.z); const Mat3 s = skew(v);
    a.m[0][0] = p.w;
    for (int j = 0; j < 3; ++j) { a.m[0][1 + j] = -v[j]; a.m[1 + j][0] = v[j]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a.m[1 + i][1 + j] = (i == j ? p.w : 0.0) - s.m[i][j];
    return a;
}
inline Mat3 bottom_right(const Mat4& a) { Mat3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[1 + i][1 + j]; return r; }
inline Mat4 operator*(const Mat4& a, const Mat4& b) {
    Mat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; }
    return r;
}

// ---------------------------------------------------------------------------------------
// Midpoint propagation, preintegration.cpp:30-127.  noise4 = {ACC_N, GYR_N, ACC_W, GYR_W}
// (preintegration.cpp:15-28 builds the 18x18 block-diagonal noise from their squares).
// ---------------------------------------------------------------------------------------
struct PreintState {
    Preint p;
    V3d acc0, gyr0;
    bool started;
};
inline void preint_reset(PreintState& s, const V3d& ba, const V3d& bg) {
    s.p.dp = V3d(); s.p.dv = V3d(); s.p.dq = Qd(); s.p.lin_ba = ba; s.p.lin_bg = bg; s.p.sum_dt = 0.0;
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) { s.p.jac[i][j] = (i == j) ? 1.0 : 0.0; s.p.cov[i][j] = 0.0; }
    s.started = false;
}

inline void preint_propagate(PreintState& s, double dt, const V3d& acc1, const V3d& gyr1, const double noise4[4]) {
    Preint& p = s.p;
    const V3d& acc0 = s.acc0; const V3d& gyr0 = s.gyr0;
    // :40-48 state
    const V3d un_acc_0 = eig_rotate(p.dq, acc0 - p.lin_ba);
    const V3d un_gyr = (gyr0 + gyr1) * 0.5 - p.lin_bg;
    const Qd res_q = eig_mul(p.dq, Qd(un_gyr.x * dt / 2, un_gyr.y * dt / 2, un_gyr.z * dt / 2, 1.0));
    const V3d un_acc_1 = eig_rotate(res_q, acc1 - p.lin_ba);
    const V3d un_acc = (un_acc_0 + un_acc_1) * 0.5;
    const V3d res_p = p.dp + p.dv * dt + un_acc * (0.5 * dt * dt);
    const V3d res_v = p.dv + un_acc * dt;

    // :50-98 F, V
    const V3d w_x = (gyr0 + gyr1) * 0.5 - p.lin_bg;
    const V3d a_0_x = acc0 - p.lin_ba;
    const V3d a_1_x = acc1 - p.lin_ba;
    const Mat3 R_w_x = skew(w_x), R_a_0_x = skew(a_0_x), R_a_1_x = skew(a_1_x);
    const Mat3 I = Mat3::identity();
    const Mat3 Rd = eig_matrix(p.dq);
    const Mat3 Rr = eig_matrix(res_q);
    const Mat3 ImW = I - R_w_x * dt;

    static thread_local double F[15][15], V[15][18];
    std::memset(F, 0, sizeof(F)); std::memset(V, 0, sizeof(V));
    auto setF = [&](int r, int c, const Mat3& b) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) F[r + i][c + j] = b.m[i][j]; };
    auto setV = [&](int r, int c, const Mat3& b) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[r + i][c + j] = b.m[i][j]; };
    setF(0, 0, I);
    setF(0, 3, (Rd * R_a_0_x) * (-0.25 * dt * dt) + ((Rr * R_a_1_x) * ImW) * (-0.25 * dt * dt));
    setF(0, 6, I * dt);
    setF(0, 9, (Rd + Rr) * (-0.25 * dt * dt));
    setF(0, 12, (Rr * R_a_1_x) * (-0.25 * dt * dt * -dt));
    setF(3, 3, ImW);
    setF(3, 12, I * (-dt));
    setF(6, 3, (Rd * R_a_0_x) * (-0.5 * dt) + ((Rr * R_a_1_x) * ImW) * (-0.5 * dt));
    setF(6, 6, I);
    setF(6, 9, (Rd + Rr) * (-0.5 * dt));
    setF(6, 12, (Rr * R_a_1_x) * (-0.5 * dt * -dt));
    setF(9, 9, I);
    setF(12, 12, I);

    const Mat3 v03 = (Rr * R_a_1_x) * (-0.25 * dt * dt * 0.5 * dt);
    const Mat3 v63 = (Rr * R_a_1_x) * (-0.5 * dt * 0.5 * dt);
    setV(0, 0, Rd * (0.25 * dt * dt));
    setV(0, 3, v03);
    setV(0, 6, Rr * (0.25 * dt * dt));
    setV(0, 9, v03);
    setV(3, 3, I * (0.5 * dt));
    setV(3, 9, I * (0.5 * dt));
    setV(6, 0, Rd * (0.5 * dt));
    setV(6, 3, v63);
    setV(6, 6, Rr * (0.5 * dt));
    setV(6, 9, v63);
    setV(9, 12, I * dt);
    setV(12, 15, I * dt);

    double nz[18];
    for (int i = 0; i < 3; ++i) {
        nz[i] = noise4[0] * noise4[0]; nz[3 + i] = noise4[1] * noise4[1];
        nz[6 + i] = noise4[0] * noise4[0]; nz[9 + i] = noise4[1] * noise4[1];
        nz[12 + i] = noise4[2] * noise4[2]; nz[15 + i] = noise4[3] * noise4[3];
    }
    // :100-101 jacobian = F*jacobian ; covariance = F cov F^T + V noise V^T
    static thread_local double tmp[15][15], tmp2[15][15];
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) { double a = 0; for (int k = 0; k < 15; ++k) a += F[i][k] * p.jac[k][j]; tmp[i][j] = a; }
    std::memcpy(p.jac, tmp, sizeof(tmp));
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) { double a = 0; for (int k = 0; k < 15; ++k) a += F[i][k] * p.cov[k][j]; tmp[i][j] = a; }
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) {
        double a = 0; for (int k = 0; k < 15; ++k) a += tmp[i][k] * F[j][k];
        double b = 0; for (int k = 0; k < 18; ++k) b += V[i][k] * nz[k] * V[j][k];
        tmp2[i][j] = a + b;
    }
    std::memcpy(p.cov, tmp2, sizeof(tmp2));

    // :117-126 commit, normalise delta_q, advance
    p.dp = res_p; p.dv = res_v;
    const double n = std::sqrt(res_q.x * res_q.x + res_q.y * res_q.y + res_q.z * res_q.z + res_q.w * res_q.w);
    p.dq = Qd(res_q.x / n, res_q.y / n, res_q.z / n, res_q.w / n);
    p.sum_dt += dt;
    s.acc0 = acc1; s.gyr0 = gyr1;
}

// preintegration.h:27-40 Append semantic: the first call seeds acc0/gyr0 with (acc0_, gyr0_).
inline void preint_append(PreintState& s, double dt, const V3d& acc, const V3d& gyr,
                          const V3d& acc0_, const V3d& gyr0_, const double noise4[4]) {
    if (!s.started) { s.acc0 = acc0_; s.gyr0 = gyr0_; s.started = true; }
    preint_propagate(s, dt, acc, gyr, noise4);
}

// ---------------------------------------------------------------------------------------
// sqrt_info = LLT(covariance.inverse()).matrixL().transpose()   imu_error.hpp:32
// [upstream] Eigen: inverse() of a 15x15 = PartialPivLU solve against identity;
// LLT = Cholesky, lower.  Both restated in their plain unblocked forms.
// Returns false only if the covariance is singular or a pivot is NaN; a non-positive Cholesky pivot follows Eigen (below).
// ---------------------------------------------------------------------------------------
inline bool sqrt_information(const double cov[15][15], double U[15][15], double prior_a = -1.0, double prior_g = -1.0) {
    const int n = 15;
    double a[15][15], inv[15][15];
    std::memcpy(a, cov, sizeof(a));
    int piv[15];
    for (int i = 0; i < n; ++i) piv[i] = i;
    for (int k = 0; k < n; ++k) {
        int best = k; double bv = std::fabs(a[k][k]);
        for (int i = k + 1; i < n; ++i) if (std::fabs(a[i][k]) > bv) { bv = std::fabs(a[i][k]); best = i; }
        if (bv == 0.0) return false;
        if (best != k) { for (int j = 0; j < n; ++j) std::swap(a[k][j], a[best][j]); std::swap(piv[k], piv[best]); }
        for (int i = k + 1; i < n; ++i) {
            a[i][k] /= a[k][k];
            const double f = a[i][k];
            for (int j = k + 1; j < n; ++j) a[i][j] -= f * a[k][j];
        }
    }
    for (int c = 0; c < n; ++c) {  // solve A x = e_c
        double y[15];
        for (int i = 0; i < n; ++i) { double s = (piv[i] == c) ? 1.0 : 0.0; for (int k = 0; k < i; ++k) s -= a[i][k] * y[k]; y[i] = s; }
        for (int i = n - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < n; ++k) s -= a[i][k] * inv[k][c]; inv[i][c] = s / a[i][i]; }
    }
    if (prior_a >= 0.0 && prior_g >= 0.0) {   // imu_error.hpp:147-149
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { inv[9 + i][9 + j] = (i == j) ? prior_a : 0.0; inv[12 + i][12 + j] = (i == j) ? prior_g : 0.0; }
    }
    double L[15][15];
    std::memset(L, 0, sizeof(L));
    for (int j = 0; j < n; ++j) {
        double d = inv[j][j];
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (d != d) return false;
        if (d <= 0.0) {
            // Eigen's llt_inplace::unblocked returns here (info() = NumericalIssue, never read by the reference) and
            // matrixL() shows the untouched lower triangle of the input from this column on.  ImuInitError gets here with
            // the reference's own priors (initializer.cpp:62: 1e4 / 1e2 over the bias blocks make cov^-1 indefinite).
            for (int c = j; c < n; ++c) for (int i = c; i < n; ++i) L[i][c] = inv[i][c];
            break;
        }
        L[j][j] = std::sqrt(d);
        for (int i = j + 1; i < n; ++i) {
            double s = inv[i][j];
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
            L[i][j] = s / L[j][j];
        }
    }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) U[i][j] = L[j][i];
    return true;
}

// preintegration.cpp:144-165 Preintegration::Evaluate (un-whitened 15-residual)
inline void imu_raw_residual(const Preint& pre, const V3d& Pi, const Qd& Qi, const V3d& Vi, const V3d& Bai, const V3d& Bgi,
                             const V3d& Pj, const Qd& Qj, const V3d& Vj, const V3d& Baj, const V3d& Bgj, double* r) {
    const Mat3 dp_dba = block3(pre.jac, 0, 9), dp_dbg = block3(pre.jac, 0, 12), dq_dbg = block3(pre.jac, 3, 12);
    const Mat3 dv_dba = block3(pre.jac, 6, 9), dv_dbg = block3(pre.jac, 6, 12);
    const V3d g(kGravity[0], kGravity[1], kGravity[2]);
    const V3d dba = Bai - pre.lin_ba, dbg = Bgi - pre.lin_bg;
    const Qd cq = eig_mul(pre.dq, q_delta(dq_dbg * dbg));
    const V3d cv = pre.dv + dv_dba * dba + dv_dbg * dbg;
    const V3d cp = pre.dp + dp_dba * dba + dp_dbg * dbg;
    const Qd Qi_inv = eig_inverse(Qi);
    const double dt = pre.sum_dt;
    const V3d rp = eig_rotate(Qi_inv, ((g * 0.5) * dt) * dt + Pj - Pi - Vi * dt) - cp;
    const Qd qe = eig_mul(eig_inverse(cq), eig_mul(Qi_inv, Qj));
    const V3d rv = eig_rotate(Qi_inv, g * dt + Vj - Vi) - cv;
    r[0] = rp.x; r[1] = rp.y; r[2] = rp.z;
    r[3] = 2 * qe.x; r[4] = 2 * qe.y; r[5] = 2 * qe.z;
    r[6] = rv.x; r[7] = rv.y; r[8] = rv.z;
    r[9] = Baj.x - Bai.x; r[10] = Baj.y - Bai.y; r[11] = Baj.z - Bai.z;
    r[12] = Bgj.x - Bgi.x; r[13] = Bgj.y - Bgi.y; r[14] = Bgj.z - Bgi.z;
}

// imu_error.hpp:17-113 ImuError::Evaluate.  params = 8 blocks (7,3,3,3,7,3,3,3); J[k] row-major
// 15 x size_k or nullptr.  Returns false when sqrt_information fails.
inline bool imu_error_evaluate(const Preint& pre, const double* const* prm, double* res, double* const* J) {
    const Qd Qi(prm[0][0], prm[0][1], prm[0][2], prm[0][3]);
    const V3d Pi(prm[0][4], prm[0][5], prm[0][6]);
    const V3d Vi(prm[1][0], prm[1][1], prm[1][2]), Bai(prm[2][0], prm[2][1], prm[2][2]), Bgi(prm[3][0], prm[3][1], prm[3][2]);
    const Qd Qj(prm[4][0], prm[4][1], prm[4][2], prm[4][3]);
    const V3d Pj(prm[4][4], prm[4][5], prm[4][6]);
    const V3d Vj(prm[5][0], prm[5][1], prm[5][2]), Baj(prm[6][0], prm[6][1], prm[6][2]), Bgj(prm[7][0], prm[7][1], prm[7][2]);

    double raw[15];
    imu_raw_residual(pre, Pi, Qi, Vi, Bai, Bgi, Pj, Qj, Vj, Baj, Bgj, raw);
    double U[15][15];
    if (!sqrt_information(pre.cov, U, pre.prior_a, pre.prior_g)) return false;
    for (int i = 0; i < 15; ++i) { double s = 0; for (int k = 0; k < 15; ++k) s += U[i][k] * raw[k]; res[i] = s; }
    if (!J) return true;

    const double dt = pre.sum_dt;
    const Mat3 dp_dba = block3(pre.jac, 0, 9), dp_dbg = block3(pre.jac, 0, 12), dq_dbg = block3(pre.jac, 3, 12);
    const Mat3 dv_dba = block3(pre.jac, 6, 9), dv_dbg = block3(pre.jac, 6, 12);
    const V3d g(kGravity[0], kGravity[1], kGravity[2]);
    const Qd Qi_inv = eig_inverse(Qi), Qj_inv = eig_inverse(Qj);
    const Mat3 Ri_inv = eig_matrix(Qi_inv);
    const Qd cq = eig_mul(pre.dq, q_delta(dq_dbg * (Bgi - pre.lin_bg)));
    const Mat3 I = Mat3::identity();

    auto whiten = [&](const double* raw_block, int cols, double* out) {
        for (int i = 0; i < 15; ++i) for (int j = 0; j < cols; ++j) {
            double s = 0; for (int k = 0; k < 15; ++k) s += U[i][k] * raw_block[k * cols + j]; out[i * cols + j] = s; }
    };
    auto put = [](double* m, int cols, int r, int c, const Mat3& b) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[(r + i) * cols + c + j] = b.m[i][j]; };

    double buf[15 * 7];
    if (J[0]) {  // :43-53 pose_i
        std::memset(buf, 0, sizeof(buf));
        put(buf, 7, 0, 4, -Ri_inv);
        put(buf, 7, 0, 0, skew(eig_rotate(Qi_inv, ((g * 0.5) * dt) * dt + Pj - Pi - Vi * dt)));
        put(buf, 7, 3, 0, -bottom_right(q_left(eig_mul(Qj_inv, Qi)) * q_right(cq)));
        put(buf, 7, 6, 0, skew(eig_rotate(Qi_inv, g * dt + Vj - Vi)));
        whiten(buf, 7, J[0]);
    }
    if (J[1]) {  // :54-61 v_i
        std::memset(buf, 0, sizeof(buf));
        put(buf, 3, 0, 0, -(Ri_inv * dt));
        put(buf, 3, 6, 0, -Ri_inv);
        whiten(buf, 3, J[1]);
    }
    if (J[2]) {  // :62-70 ba_i
        std::memset(buf, 0, sizeof(buf));
        put(buf, 3, 0, 0, -dp_dba);
        put(buf, 3, 6, 0, -dv_dba);
        put(buf, 3, 9, 0, -I);
        whiten(buf, 3, J[2]);
    }
    if (J[3]) {  // :71-80 bg_i
        std::memset(buf, 0, sizeof(buf));
        put(buf, 3, 0, 0, -dp_dbg);
        put(buf, 3, 3, 0, -(bottom_right(q_left(eig_mul(eig_mul(Qj_inv, Qi), pre.dq))) * dq_dbg));
        put(buf, 3, 6, 0, -dv_dbg);
        put(buf, 3, 12, 0, -I);
        whiten(buf, 3, J[3]);
    }
    if (J[4]) {  // :81-89 pose_j
        std::memset(buf, 0, sizeof(buf));
        put(buf, 7, 0, 4, Ri_inv);
        put(buf, 7, 3, 0, bottom_right(q_left(eig_mul(eig_mul(eig_inverse(cq), Qi_inv), Qj))));
        whiten(buf, 7, J[4]);
    }
    if (J[5]) {  // :90-96 v_j
        std::memset(buf, 0, sizeof(buf));
        put(buf, 3, 6, 0, Ri_inv);
        whiten(buf, 3, J[5]);
    }
    if (J[6]) {  // :97-103 ba_j
        std::memset(buf, 0, sizeof(buf));
        put(buf, 3, 9, 0, I);
        whiten(buf, 3, J[6]);
    }
    if (J[7]) {  // :104-110 bg_j
        std::memset(buf, 0, sizeof(buf));
        put(buf, 3, 12, 0, I);
        whiten(buf, 3, J[7]);
    }
    return true;
}

}  // namespace oracle
