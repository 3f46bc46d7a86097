// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the REFERENCE's own factor functors (the headers are compiled where they lie under /root/reference, never copied)
// on inputs from a case file and writes residuals + Jacobians.  Third parties are replaced by the stand-ins in
// oracle/ref_compat (Eigen/Sophus/OpenCV types, <ceres/*.h>); derivatives come from instantiating the functors' operator()
// with the oracle's dual numbers, which is what ceres::AutoDiffCostFunction does.  Built by `make -C oracle ref` into
// oracle/_ref/ref_factors (git-ignored); used by tests/golden/make_ref_golden.py to produce the committed fixture
// tests/golden/ref_factors.npz that pins the oracle (and the CUDA path) for the visual, lidar and pose factor kinds.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lvio_fusion/ceres/visual_error.hpp"
#include "lvio_fusion/ceres/lidar_error.hpp"
#include "lvio_fusion/ceres/pose_error.hpp"
#include "lvio_fusion/ceres/imu_error.hpp"

namespace lvio_fusion {
std::vector<Camera::Ptr> Camera::devices_;      // defined in the reference's src/visual/camera.cpp, which is not compiled here
double Camera::baseline = 1;
std::vector<Imu::Ptr> Imu::devices_;            // src/imu/imu.cpp, not compiled here
}
using namespace lvio_fusion;
using oracle::Dual;

static std::vector<double> in;
static size_t pos = 0;
static double next() { if (pos >= in.size()) { fprintf(stderr, "case file too short\n"); exit(2); } return in[pos++]; }
static void take(double* dst, int n) { for (int i = 0; i < n; ++i) dst[i] = next(); }
static FILE* out;
static void put(const double* v, int n) { fwrite(v, sizeof(double), n, out); }

// evaluate functor f at the blocks x (sizes sz) with N = sum(sz) duals; writes r[R] then J[R x N] row-major
template <int R, int N, class F, class Call>
static void eval(const F& f, const double* x, Call call) {
    Dual<N> X[N], Y[R];
    for (int k = 0; k < N; ++k) X[k] = Dual<N>::seed(x[k], k);
    call(f, X, Y);
    double r[R], J[R * N];
    for (int i = 0; i < R; ++i) { r[i] = Y[i].v; for (int k = 0; k < N; ++k) J[i * N + k] = Y[i].d[k]; }
    put(r, R); put(J, R * N);
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_factors <cases.bin> <out.bin>\n"); return 64; }
    FILE* fi = fopen(argv[1], "rb"); if (!fi) return 2;
    fseek(fi, 0, SEEK_END); const long bytes = ftell(fi); fseek(fi, 0, SEEK_SET);
    in.resize(bytes / sizeof(double));
    if (fread(in.data(), sizeof(double), in.size(), fi) != in.size()) return 2;
    fclose(fi);
    out = fopen(argv[2], "wb"); if (!out) return 2;
    // cameras: fx fy cx cy extrinsic[7], twice (Camera::Get(0) = left / cam0, Camera::Get(1) = right / cam1)
    double cam[22]; take(cam, 22);
    for (int c = 0; c < 2; ++c) Camera::Create(cam[11 * c], cam[11 * c + 1], cam[11 * c + 2], cam[11 * c + 3], SE3d(cam + 11 * c + 4));
    Camera::Ptr left = Camera::Get(0), right = Camera::Get(1);
    const int n_tf = (int)next(), n_po = (int)next(), n_tc = (int)next(), n_lidar = (int)next(), n_pg = (int)next(), n_pe = (int)next(), n_pr = (int)next(), n_imu = (int)next();
    for (int i = 0; i < n_tf; ++i) {           // a1: first_ob(2) ob(2) w | rho T1 T2
        double c[5], x[15]; take(c, 5); take(x, 15);
        TwoFrameReprojectionError f(Vector2d(c[0], c[1]), Vector2d(c[2], c[3]), left, right, c[4]);
        eval<2, 15>(f, x, [](const TwoFrameReprojectionError& g, const Dual<15>* X, Dual<15>* Y) { g(X, X + 1, X + 8, Y); });
    }
    for (int i = 0; i < n_po; ++i) {           // a2: ob(2) pw(3) w | T
        double c[6], x[7]; take(c, 6); take(x, 7);
        PoseOnlyReprojectionError f(Vector2d(c[0], c[1]), Vector3d(c[2], c[3], c[4]), left, c[5]);
        eval<2, 7>(f, x, [](const PoseOnlyReprojectionError& g, const Dual<7>* X, Dual<7>* Y) { g(X, Y); });
    }
    for (int i = 0; i < n_tc; ++i) {           // a3: left_ob(2) right_ob(2) w | rho
        double c[5], x[1]; take(c, 5); take(x, 1);
        TwoCameraReprojectionError f(Vector2d(c[0], c[1]), Vector2d(c[2], c[3]), left, right, c[4]);
        eval<2, 1>(f, x, [](const TwoCameraReprojectionError& g, const Dual<1>* X, Dual<1>* Y) { g(X, Y); });
    }
    for (int i = 0; i < n_lidar; ++i) {        // a5: mode p pa pb pc Twc1 rpyxyz w ; the three free scalars are read from rpyxyz
        const int mode = (int)next();
        double p[3], pa[3], pb[3], pc[3], T[7], e[6]; take(p, 3); take(pa, 3); take(pb, 3); take(pc, 3); take(T, 7); take(e, 6);
        const double w = next();
        LidarPlaneError base(Vector3d(p[0], p[1], p[2]), Vector3d(pa[0], pa[1], pa[2]), Vector3d(pb[0], pb[1], pb[2]), Vector3d(pc[0], pc[1], pc[2]));
        if (mode == 0) {
            LidarPlaneErrorRPZ f(base, SE3d(T), e, w);
            const double x[3] = {e[1], e[2], e[5]};
            eval<1, 3>(f, x, [](const LidarPlaneErrorRPZ& g, const Dual<3>* X, Dual<3>* Y) { g(X, X + 1, X + 2, Y); });
        } else {
            LidarPlaneErrorYXY f(base, SE3d(T), e, w);
            const double x[3] = {e[0], e[3], e[4]};
            eval<1, 3>(f, x, [](const LidarPlaneErrorYXY& g, const Dual<3>* X, Dual<3>* Y) { g(X, X + 1, X + 2, Y); });
        }
    }
    for (int i = 0; i < n_pg; ++i) {           // a6 PoseGraphError: last(7) pose(7) w v | T1 T2 ; also emits the stored rpyxyz_
        double a[7], b[7], x[14]; take(a, 7); take(b, 7); const double w = next(), v = next(); take(x, 14);
        PoseGraphError f(SE3d(a), SE3d(b), w, v);
        eval<6, 14>(f, x, [](const PoseGraphError& g, const Dual<14>* X, Dual<14>* Y) { g(X, X + 7, Y); });
    }
    for (int i = 0; i < n_pe; ++i) {           // a6 PoseError: pose(7) w v | T
        double a[7], x[7]; take(a, 7); const double w = next(), v = next(); take(x, 7);
        PoseError f(SE3d(a), w, v);
        eval<6, 7>(f, x, [](const PoseError& g, const Dual<7>* X, Dual<7>* Y) { g(X, Y); });
    }
    for (int i = 0; i < n_pr; ++i) {           // a6 PoseErrorRPZ / PoseErrorYXY: mode rpyxyz(6) w | 3 scalars
        const int mode = (int)next();
        double e[6], x[3]; take(e, 6); const double w = next(); take(x, 3);
        if (mode == 0) { PoseErrorRPZ f(e, w); eval<3, 3>(f, x, [](const PoseErrorRPZ& g, const Dual<3>* X, Dual<3>* Y) { g(X, X + 1, X + 2, Y); }); }
        else { PoseErrorYXY f(e, w); eval<3, 3>(f, x, [](const PoseErrorYXY& g, const Dual<3>* X, Dual<3>* Y) { g(X, X + 1, X + 2, Y); }); }
    }
    // a4 ImuError (imu_error.hpp:12-122) on top of Preintegration::Append / Propagate (preintegration.h:27-40,
    // src/preintegration.cpp:30-127, compiled in place): noise4 once, then per case  ba bg acc0 gyr0 n_samples (dt acc gyr)*
    // and the 8 parameter blocks pose_i v_i ba_i bg_i pose_j v_j ba_j bg_j.  Emits the preintegration record (delta_p,
    // delta_q xyzw, delta_v, linearized ba / bg, sum_dt, jacobian, covariance), then r[15] and the 8 row-major Jacobians.
    if (n_imu > 0) {
        double nz[4]; take(nz, 4);                     // ACC_N GYR_N ACC_W GYR_W
        Imu::Create(SE3d(), nz[0], nz[2], nz[1], nz[3], 9.81007);
    }
    for (int i = 0; i < n_imu; ++i) {
        double ba[3], bg[3], a0[3], g0[3]; take(ba, 3); take(bg, 3); take(a0, 3); take(g0, 3);
        const int ns = (int)next();
        imu::Preintegration::Ptr pre = imu::Preintegration::Create(Bias(Vector3d(ba[0], ba[1], ba[2]), Vector3d(bg[0], bg[1], bg[2])));
        for (int k = 0; k < ns; ++k) {
            double smp[7]; take(smp, 7);
            pre->Append(smp[0], Vector3d(smp[1], smp[2], smp[3]), Vector3d(smp[4], smp[5], smp[6]), Vector3d(a0[0], a0[1], a0[2]), Vector3d(g0[0], g0[1], g0[2]));
        }
        double rec[17] = {pre->delta_p.x(), pre->delta_p.y(), pre->delta_p.z(), pre->delta_q.x(), pre->delta_q.y(), pre->delta_q.z(), pre->delta_q.w(),
                          pre->delta_v.x(), pre->delta_v.y(), pre->delta_v.z(), pre->linearized_ba.x(), pre->linearized_ba.y(), pre->linearized_ba.z(),
                          pre->linearized_bg.x(), pre->linearized_bg.y(), pre->linearized_bg.z(), pre->sum_dt};
        put(rec, 17); put(pre->jacobian.data(), 225); put(pre->covariance.data(), 225);
        double x[32]; take(x, 32);
        const double* prm[8] = {x, x + 7, x + 10, x + 13, x + 16, x + 23, x + 26, x + 29};
        double r[15], J0[105], J1[45], J2[45], J3[45], J4[105], J5[45], J6[45], J7[45];
        double* J[8] = {J0, J1, J2, J3, J4, J5, J6, J7};
        ImuError f(pre);
        f.Evaluate(prm, r, J);
        put(r, 15); put(J0, 105); put(J1, 45); put(J2, 45); put(J3, 45); put(J4, 105); put(J5, 45); put(J6, 45); put(J7, 45);
        // ImuInitError (imu_error.hpp:124-229, imu::FullBA): same preintegration, priors 1e8 / 1e8, blocks pose_i v_i ba_i bg_i pose_j v_j
        ImuInitError fi(pre, 1e8, 1e8);
        const double* prm6[6] = {x, x + 7, x + 10, x + 13, x + 16, x + 23};
        double* J6b[6] = {J0, J1, J2, J3, J4, J5};
        fi.Evaluate(prm6, r, J6b);
        put(r, 15); put(J0, 105); put(J1, 45); put(J2, 45); put(J3, 45); put(J4, 105); put(J5, 45);
        // ... and with the priors Initializer::Initialize really passes (src/initializer.cpp:62: 1e4 / 1e2): written over the
        // bias blocks they leave cov^-1 indefinite, so this is the LLT-returns-early path (see mini_eigen.h LLT)
        ImuInitError fr(pre, 1e4, 1e2);
        fr.Evaluate(prm6, r, J6b);
        put(r, 15); put(J0, 105); put(J1, 45); put(J2, 45); put(J3, 45); put(J4, 105); put(J5, 45);
    }
    fclose(out);
    return 0;
}
