// tests/cpp/test_host_solver.cpp -- host-side small solves through the ceres-shaped shim, no GPU involved.
//   posegraph <in.bin> <out.bin> : pose chain with PoseGraphError edges (factors.h) + generic AutoDiff translation priors,
//                                  the shape of Navsat::OptimizeAB (navsat.cpp:273-306); result compared with the oracle LM
//   navsat <seed>                 : yaw/x/y alignment in two stages with constant / variable blocks and a bounded block
//                                  (navsat.cpp:104-129,245-246), functor written against ceres::QuaternionRotatePoint
//   autodiff                      : AutoDiffCostFunction Jacobians vs central differences
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "lvio_b200/factors.h"

struct SE3 { double d[7]; const double* data() const { return d; } double* data() { return d; } };

struct TranslationPrior {            // TError-shaped: residual = w (t - t0)
    double t0[3], w;
    template <typename T> bool operator()(const T* pose, T* r) const { for (int i = 0; i < 3; ++i) r[i] = T(w) * (pose[4 + i] - T(t0[i])); return true; }
};

struct YawXYAlign {                  // NavsatInitError-shaped: p0 ~ Rz(yaw) p1 + (x, y, 0)
    double p0[3], p1[3], s[3];
    template <typename T> bool operator()(const T* yaw, const T* x, const T* y, T* r) const {
        const T half = yaw[0] * T(0.5);
        const T q[4] = {cos(half), T(0), T(0), sin(half)};          // w-first, rotation about z
        const T p[3] = {T(p1[0]), T(p1[1]), T(p1[2])};
        T out[3];
        ceres::QuaternionRotatePoint(q, p, out);
        r[0] = T(s[0]) * (T(p0[0]) - (out[0] + x[0])); r[1] = T(s[1]) * (T(p0[1]) - (out[1] + y[0])); r[2] = T(s[2]) * (T(p0[2]) - out[2]);
        return true;
    }
};

static double urand(unsigned& st) { st = st * 1664525u + 1013904223u; return (double)(st >> 8) / 16777216.0; }

static int run_posegraph(const char* in, const char* out) {
    FILE* f = fopen(in, "rb"); if (!f) return 2;
    int hdr[3]; if (fread(hdr, sizeof(int), 3, f) != 3) return 2;
    const int n = hdr[0], max_iter = hdr[2];
    std::vector<SE3> pose(n), meas(n); std::vector<double> gps(3 * (size_t)n); double w[3];
    if (fread(pose.data(), sizeof(SE3), n, f) != (size_t)n || fread(meas.data(), sizeof(SE3), n, f) != (size_t)n || fread(gps.data(), sizeof(double), 3 * n, f) != 3 * (size_t)n ||
        fread(w, sizeof(double), 3, f) != 3) return 2;
    fclose(f);
    ceres::Problem problem;
    ceres::LossFunction* loss = hdr[1] ? new ceres::HuberLoss(0.5) : nullptr;
    ceres::LocalParameterization* lp = new ceres::ProductParameterization(new ceres::EigenQuaternionParameterization(), new ceres::IdentityParameterization(3));
    for (int i = 0; i < n; ++i) problem.AddParameterBlock(pose[i].data(), 7, lp);
    for (int i = 0; i + 1 < n; ++i) problem.AddResidualBlock(lvio_fusion::PoseGraphError::Create(meas[i], meas[i + 1], w[0], w[1]), nullptr, pose[i].data(), pose[i + 1].data());
    for (int i = 0; i < n; ++i) {
        TranslationPrior* tp = new TranslationPrior(); for (int k = 0; k < 3; ++k) tp->t0[k] = gps[3 * i + k]; tp->w = w[2];
        problem.AddResidualBlock(new ceres::AutoDiffCostFunction<TranslationPrior, 3, 7>(tp), loss, pose[i].data());
    }
    ceres::Solver::Options options; options.linear_solver_type = ceres::DENSE_QR; options.max_num_iterations = max_iter;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    FILE* o = fopen(out, "wb"); if (!o) return 2;
    const double head[4] = {summary.initial_cost, summary.final_cost, (double)(summary.num_successful_steps + summary.num_unsuccessful_steps), (double)summary.termination_type};
    fwrite(head, sizeof(double), 4, o); fwrite(pose.data(), sizeof(SE3), n, o); fclose(o);
    fprintf(stderr, "%s\n", summary.message.c_str());
    return summary.termination_type == ceres::FAILURE ? 3 : 0;
}

static int run_navsat(unsigned seed) {
    const double yaw_true = 0.7, x_true = 12.5, y_true = -4.25;
    double para[6] = {0.0, 0, 0, 0.0, 0.0, 0};      // yaw .. x y
    ceres::Problem problem;
    problem.AddParameterBlock(para, 1); problem.AddParameterBlock(para + 3, 1); problem.AddParameterBlock(para + 4, 1);
    problem.SetParameterBlockConstant(para + 3); problem.SetParameterBlockConstant(para + 4);
    for (int i = 0; i < 40; ++i) {
        YawXYAlign* fct = new YawXYAlign();
        fct->p1[0] = 50 * urand(seed) - 25; fct->p1[1] = 50 * urand(seed) - 25; fct->p1[2] = urand(seed);
        fct->p0[0] = std::cos(yaw_true) * fct->p1[0] - std::sin(yaw_true) * fct->p1[1] + x_true + 0.01 * (urand(seed) - 0.5);
        fct->p0[1] = std::sin(yaw_true) * fct->p1[0] + std::cos(yaw_true) * fct->p1[1] + y_true + 0.01 * (urand(seed) - 0.5);
        fct->p0[2] = fct->p1[2];
        fct->s[0] = fct->s[1] = 10.0; fct->s[2] = 1.0;
        problem.AddResidualBlock(new ceres::AutoDiffCostFunction<YawXYAlign, 3, 1, 1, 1>(fct), nullptr, para, para + 3, para + 4);
    }
    ceres::Solver::Options options; options.linear_solver_type = ceres::DENSE_QR;
    ceres::Solver::Summary s1, s2, s3;
    ceres::Solve(options, &problem, &s1);                       // yaw only (navsat.cpp:109-125)
    const double yaw_stage1 = para[0], x_stage1 = para[3];
    problem.SetParameterBlockVariable(para + 3); problem.SetParameterBlockVariable(para + 4);
    ceres::Solve(options, &problem, &s2);                       // all three (navsat.cpp:127-129)
    printf("stage1 yaw %.12f x %.12f cost %.9e -> %.9e term %d\n", yaw_stage1, x_stage1, s1.initial_cost, s1.final_cost, (int)s1.termination_type);
    printf("stage2 yaw %.12f x %.12f y %.12f cost %.9e term %d\n", para[0], para[3], para[4], s2.final_cost, (int)s2.termination_type);
    // bounded block (navsat.cpp:245-246): x may not leave [x_true + 1, x_true + 3]
    para[0] = 0; para[3] = x_true + 2; para[4] = 0;
    problem.SetParameterLowerBound(para + 3, 0, x_true + 1.0); problem.SetParameterUpperBound(para + 3, 0, x_true + 3.0);
    ceres::Solve(options, &problem, &s3);
    printf("bounded x %.12f y %.12f term %d\n", para[3], para[4], (int)s3.termination_type);
    return 0;
}

static int run_autodiff() {
    lvb::host::PoseGraphFunctor* f = new lvb::host::PoseGraphFunctor();
    const double e[6] = {0.1, -0.05, 0.02, 1.0, 0.2, -0.1}; for (int i = 0; i < 6; ++i) f->e[i] = e[i]; f->w = 3.0; f->v = 0.7;
    ceres::AutoDiffCostFunction<lvb::host::PoseGraphFunctor, 6, 7, 7> cost(f);
    double a[7] = {0.05, -0.1, 0.2, 0.97, 1.0, 2.0, -0.5}, b[7] = {-0.02, 0.15, 0.1, 0.98, 2.1, 2.3, -0.2};
    const double* p[2] = {a, b};
    double r[6], Ja[42], Jb[42]; double* J[2] = {Ja, Jb};
    if (!cost.Evaluate(p, r, J)) return 2;
    double worst = 0;
    for (int blk = 0; blk < 2; ++blk) for (int k = 0; k < 7; ++k) {
        double* x = blk ? b : a; const double keep = x[k], h = 1e-6;
        double rp[6], rm[6];
        x[k] = keep + h; cost.Evaluate(p, rp, nullptr); x[k] = keep - h; cost.Evaluate(p, rm, nullptr); x[k] = keep;
        for (int i = 0; i < 6; ++i) { const double fd = (rp[i] - rm[i]) / (2 * h), an = (blk ? Jb : Ja)[i * 7 + k]; worst = std::fmax(worst, std::fabs(fd - an) / std::fmax(1.0, std::fabs(an))); }
    }
    printf("autodiff_vs_fd %.3e\n", worst);
    // rotation.h subset: rotate by a non-unit quaternion == rotate by its normalisation; product is associative with rotation
    const double q[4] = {1.9, 0.2, -0.6, 0.4}, pt[3] = {0.3, -1.2, 2.0};
    double r1[3], r2[3], qq[4], r3[3], r4[3];
    ceres::QuaternionRotatePoint(q, pt, r1);
    const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double u[4] = {q[0] / nq, q[1] / nq, q[2] / nq, q[3] / nq};
    ceres::UnitQuaternionRotatePoint(u, pt, r2);
    const double q2[4] = {0.8, -0.1, 0.3, 0.5};
    ceres::QuaternionProduct(q, q2, qq); ceres::QuaternionRotatePoint(qq, pt, r3);
    double tmp[3]; ceres::QuaternionRotatePoint(q2, pt, tmp); ceres::QuaternionRotatePoint(q, tmp, r4);
    double e1 = 0, e2 = 0, nrm = 0;
    for (int i = 0; i < 3; ++i) { e1 = std::fmax(e1, std::fabs(r1[i] - r2[i])); e2 = std::fmax(e2, std::fabs(r3[i] - r4[i])); nrm += r1[i] * r1[i]; }
    printf("rotation %.3e %.3e norm %.12f\n", e1, e2, std::sqrt(nrm));
    // NumericDiffCostFunction (FORWARD as ImuInitGError uses it, and CENTRAL) against the AutoDiff Jacobian of the same functor
    struct Plain {          // the non-template call operator NumericDiff wants, on top of the templated functor
        lvb::host::PoseGraphFunctor f;
        bool operator()(const double* x0, const double* x1, double* res) const { return f(x0, x1, res); }
    };
    Plain* pf = new Plain(); Plain* pc = new Plain();
    for (int i = 0; i < 6; ++i) { pf->f.e[i] = e[i]; pc->f.e[i] = e[i]; } pf->f.w = pc->f.w = 3.0; pf->f.v = pc->f.v = 0.7;
    ceres::NumericDiffCostFunction<Plain, ceres::FORWARD, 6, 7, 7> fwd(pf);
    ceres::NumericDiffCostFunction<Plain, ceres::CENTRAL, 6, 7, 7> cen(pc);
    double rn[6], Jfa[42], Jfb[42], Jca[42], Jcb[42]; double* Jf[2] = {Jfa, Jfb}; double* Jc[2] = {Jca, Jcb};
    if (!fwd.Evaluate(p, rn, Jf) || !cen.Evaluate(p, rn, Jc)) return 3;
    double wf = 0, wc = 0, wr = 0;
    for (int i = 0; i < 6; ++i) wr = std::fmax(wr, std::fabs(rn[i] - r[i]));
    for (int i = 0; i < 42; ++i) {
        wf = std::fmax(wf, std::fmax(std::fabs(Jfa[i] - Ja[i]), std::fabs(Jfb[i] - Jb[i])));
        wc = std::fmax(wc, std::fmax(std::fabs(Jca[i] - Ja[i]), std::fabs(Jcb[i] - Jb[i])));
    }
    printf("numericdiff %.3e %.3e %.3e\n", wf, wc, wr);
    return 0;
}

// PoseGraphError / PoseError of factors.h evaluated on the host (the evaluators the off-path solves use) on the cases of the
// reference fixture: n, then n x (last[7] pose[7] w v T1[7] T2[7]), then n x (pose[7] w v T[7]); writes r and J per case
static int run_refcases(const char* in, const char* out) {
    FILE* f = fopen(in, "rb"); if (!f) return 2;
    int n; if (fread(&n, sizeof(int), 1, f) != 1) return 2;
    std::vector<double> pg((size_t)n * 30), pe((size_t)n * 16);
    if (fread(pg.data(), sizeof(double), pg.size(), f) != pg.size() || fread(pe.data(), sizeof(double), pe.size(), f) != pe.size()) return 2;
    fclose(f);
    FILE* o = fopen(out, "wb"); if (!o) return 2;
    for (int i = 0; i < n; ++i) {
        const double* c = &pg[(size_t)i * 30];
        SE3 a, b; memcpy(a.d, c, 56); memcpy(b.d, c + 7, 56);
        ceres::CostFunction* cost = lvio_fusion::PoseGraphError::Create(a, b, c[14], c[15]);
        const double* p[2] = {c + 16, c + 23};
        double r[6], J1[42], J2[42]; double* J[2] = {J1, J2};
        if (!cost->Evaluate(p, r, J)) return 3;
        fwrite(r, sizeof(double), 6, o);
        for (int row = 0; row < 6; ++row) { fwrite(J1 + 7 * row, sizeof(double), 7, o); fwrite(J2 + 7 * row, sizeof(double), 7, o); }
        delete cost;
    }
    for (int i = 0; i < n; ++i) {
        const double* c = &pe[(size_t)i * 16];
        SE3 a; memcpy(a.d, c, 56);
        ceres::CostFunction* cost = lvio_fusion::PoseError::Create(a, c[7], c[8]);
        const double* p[1] = {c + 9};
        double r[6], J1[42]; double* J[1] = {J1};
        if (!cost->Evaluate(p, r, J)) return 3;
        fwrite(r, sizeof(double), 6, o); fwrite(J1, sizeof(double), 42, o);
        delete cost;
    }
    fclose(o);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 4 && std::string(argv[1]) == "posegraph") return run_posegraph(argv[2], argv[3]);
    if (argc >= 4 && std::string(argv[1]) == "refcases") return run_refcases(argv[2], argv[3]);
    if (argc >= 3 && std::string(argv[1]) == "navsat") return run_navsat((unsigned)atoi(argv[2]));
    if (argc >= 2 && std::string(argv[1]) == "autodiff") return run_autodiff();
    fprintf(stderr, "usage: test_host_solver posegraph in out | navsat seed | autodiff\n");
    return 64;
}
