// tests/hostcheck/hostcheck.cpp -- TEST HARNESS ONLY.
// Compiles the __host__ __device__ factor arithmetic of lvio_fusion_b200/csrc/lvb_math.cuh with
// g++ so that the closed-form Jacobians can be checked against the CPU oracle without a GPU.
// It is never loaded by the product; the kernels call the same functions on the device.
#include <cstring>
#include "../../lvio_fusion_b200/csrc/lvb_math.cuh"
using namespace lvb;

extern "C" {

static Cams load_cams(const double* cam22) { Cams K; K.c0 = make_cam(cam22); K.c1 = make_cam(cam22 + 11); return K; }

// ambient layout identical to lvb_ba_eval: r[2], J[2x15]
void hc_two_frame(const double* cam22, const double* c, double rho, const double* T1, const double* T2, double* r, double* J, double* Jt /*2x13 tangent: rho,6,6*/) {
    const Cams K = load_cams(cam22); TwoFrameLin o; UQ u1, u2;
    two_frame_lin(K, c[0], c[1], c[2], c[3], c[4], rho, T1, T2, o, &u1, &u2);
    r[0] = o.r[0]; r[1] = o.r[1];
    for (int row = 0; row < 2; ++row) { J[15 * row] = o.Jrho[row]; Jt[13 * row] = o.Jrho[row]; for (int k = 0; k < 6; ++k) { Jt[13 * row + 1 + k] = o.J1[6 * row + k]; Jt[13 * row + 7 + k] = o.J2[6 * row + k]; } }
    pose_block_to_ambient(u1, o.J1, 2, J + 1, 15);
    pose_block_to_ambient(u2, o.J2, 2, J + 8, 15);
}
void hc_pose_only(const double* cam22, const double* c, const double* T, double* r, double* J /*2x7*/, double* Jt /*2x6*/) {
    const Cams K = load_cams(cam22); PoseOnlyLin o; UQ u;
    pose_only_lin(K, c[0], c[1], v3(c[2], c[3], c[4]), c[5], T, o, &u);
    r[0] = o.r[0]; r[1] = o.r[1]; std::memcpy(Jt, o.J, sizeof(o.J));
    pose_block_to_ambient(u, o.J, 2, J, 7);
}
void hc_two_camera(const double* cam22, const double* c, double rho, double* r, double* J /*2*/) {
    const Cams K = load_cams(cam22); TwoCameraLin o;
    two_camera_lin(K, c[0], c[1], c[2], c[3], c[4], rho, o);
    r[0] = o.r[0]; r[1] = o.r[1]; J[0] = o.Jrho[0]; J[1] = o.Jrho[1];
}
void hc_pose_graph(const double* c, const double* T1, const double* T2, double* r, double* J) { pose_graph_eval(c, T1, T2, r, J); }
void hc_pose_prior(const double* c, const double* T, double* r, double* J) { pose_prior_eval(c, T, r, J); }

// c467 -> whitened residual r[15] and ambient J[15x32]
int hc_imu(const double* c467, const double* Ti, const double* Vi, const double* Bai, const double* Bgi,
           const double* Tj, const double* Vj, const double* Baj, const double* Bgj, double* r, double* J) {
    ImuConst c;
    c.dp = v3(c467[0], c467[1], c467[2]); c.dq = q4(c467[3], c467[4], c467[5], c467[6]); c.dv = v3(c467[7], c467[8], c467[9]);
    c.lin_ba = v3(c467[10], c467[11], c467[12]); c.lin_bg = v3(c467[13], c467[14], c467[15]); c.sum_dt = c467[16];
    const double* jac = c467 + 17;
    auto blk = [&](int r0, int c0) { M3 b; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) b.m[3 * i + j] = jac[(r0 + i) * 15 + c0 + j]; return b; };
    c.dp_dba = blk(0, 9); c.dp_dbg = blk(0, 12); c.dq_dbg = blk(3, 12); c.dv_dba = blk(6, 9); c.dv_dbg = blk(6, 12);
    double U[225], a[225], inv[225];
    const int rc = sqrt_information(c467 + 242, U, a, inv, c467[467], c467[468]);
    if (rc) return rc;
    double raw[15], Jr[15 * 32];
    imu_raw_residual(c, Ti, Vi, Bai, Bgi, Tj, Vj, Baj, Bgj, raw);
    for (int i = 0; i < 15 * 32; ++i) Jr[i] = 0;
    imu_raw_jacobian(c, Ti, Vi, Bgi, Tj, Vj, Jr);
    for (int i = 0; i < 15; ++i) { double s = 0; for (int k = 0; k < 15; ++k) s += U[15 * i + k] * raw[k]; r[i] = s; }
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 32; ++j) { double s = 0; for (int k = 0; k < 15; ++k) s += U[15 * i + k] * Jr[32 * k + j]; J[32 * i + j] = s; }
    return 0;
}
int hc_sqrt_information(const double* cov225, double* U225, double prior_a, double prior_g) { double a[225], inv[225]; return sqrt_information(cov225, U225, a, inv, prior_a, prior_g); }
void hc_icp_point(int mode, const double* Twc1, const double* rpyxyz, const double* c10, double* r, double* J) {
    const IcpFrame f = icp_frame(mode, Twc1, rpyxyz);
    *r = icp_point(f, v3(c10[0], c10[1], c10[2]), v3(c10[3], c10[4], c10[5]), v3(c10[6], c10[7], c10[8]), c10[9], J);
}
void hc_pose_plus(const double* x, const double* d, double* out) { pose_plus(x, d, out); }
void hc_ambient_row_to_tangent(const double* q, const double* a7, double* t6) { ambient_row_to_tangent(q, a7, t6); }
}
