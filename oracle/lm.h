// oracle/lm.h -- trust-region Levenberg-Marquardt loop of the CPU oracle
// (TEST INFRASTRUCTURE ONLY, parity unpinned).
//
// Restates the published behaviour of Ceres Solver's TrustRegionMinimizer +
// LevenbergMarquardtStrategy with default options [upstream, Ceres 2.0-2.1, NOT vendored in
// /root/reference; call sites: backend.cpp:204-211, mapping.cpp:159-164,171-176,
// backend.cpp:262-267].  Defaults: 50 iterations, radius0 1e4, min_lm_diagonal 1e-6,
// max_lm_diagonal 1e32, function_tolerance 1e-6, gradient_tolerance 1e-10,
// parameter_tolerance 1e-8, min_relative_decrease 1e-3, jacobi_scaling on,
// max_num_consecutive_invalid_steps 5, monotonic steps, no inner iterations.
//
// The model works on the un-scaled normal equations.  With Jacobi scaling s_j (fixed at the
// first linearisation, s_j = 1/(1+sqrt(H_jj))) Ceres solves (S H S + D^2) y = -S g with
// D^2_j = clamp(s_j^2 H_jj, 1e-6, 1e32)/radius and takes delta = S y; that is identical to
//     (H + Lambda) delta = -g ,  Lambda_j = clamp(s_j^2 H_jj, 1e-6, 1e32) / (radius s_j^2)
// and  model_cost_change = -(g.delta + delta.H.delta/2) = (delta.Lambda.delta - g.delta)/2.
#pragma once
#include <chrono>
#include <cmath>
#include <vector>

namespace oracle {

struct LmOptions {
    int max_num_iterations = 50;
    double max_solver_time_in_seconds = 1e6;
    double function_tolerance = 1e-6;
    double gradient_tolerance = 1e-10;
    double parameter_tolerance = 1e-8;
    double initial_trust_region_radius = 1e4;
    double max_trust_region_radius = 1e16;
    double min_trust_region_radius = 1e-32;
    double min_lm_diagonal = 1e-6;
    double max_lm_diagonal = 1e32;
    double min_relative_decrease = 1e-3;
    int jacobi_scaling = 1;
    int max_num_consecutive_invalid_steps = 5;
};

enum LmTermination { LM_CONVERGENCE = 0, LM_NO_CONVERGENCE = 1, LM_FAILURE = 2 };

struct LmSummary {
    double initial_cost = 0, final_cost = 0;
    int num_iterations = 0;        // LM steps attempted (successful + unsuccessful + invalid)
    int num_successful_steps = 0;
    int termination = LM_NO_CONVERGENCE;
    double final_radius = 0;
    double total_time_s = 0;
};

// Model concept:
//   int    dim();
//   double linearize(std::vector<double>& g, std::vector<double>& hdiag);  // at x; returns cost
//   bool   solve(const std::vector<double>& lambda, std::vector<double>& delta); // (H+L)d=-g
//   double candidate_cost(const std::vector<double>& delta);   // cost at Plus(x,delta); keeps candidate
//   void   accept();                                           // x <- candidate
//   double x_norm();  double step_norm();                      // ambient norms
//   double gradient_max_norm(const std::vector<double>& g);    // ||x - Plus(x,-g)||_inf
template <class Model>
void lm_minimize(Model& m, const LmOptions& o, LmSummary& out) {
    const auto t0 = std::chrono::steady_clock::now();
    auto elapsed = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    const int n = m.dim();
    std::vector<double> g(n), hdiag(n), scale(n, 1.0), lambda(n), delta(n);

    double x_cost = m.linearize(g, hdiag);
    out.initial_cost = x_cost;
    if (o.jacobi_scaling) for (int j = 0; j < n; ++j) scale[j] = 1.0 / (1.0 + std::sqrt(hdiag[j]));
    double radius = o.initial_trust_region_radius;
    double decrease_factor = 2.0;
    int invalid = 0;
    int iter = 0;
    bool last_successful = false;
    out.termination = LM_NO_CONVERGENCE;

    if (m.gradient_max_norm(g) <= o.gradient_tolerance) { out.termination = LM_CONVERGENCE; goto done; }

    while (true) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue
        if (elapsed() >= o.max_solver_time_in_seconds) break;
        if (iter >= o.max_num_iterations) break;
        if (last_successful && m.gradient_max_norm(g) <= o.gradient_tolerance) { out.termination = LM_CONVERGENCE; break; }
        if (radius <= o.min_trust_region_radius) { out.termination = LM_CONVERGENCE; break; }
        ++iter;
        last_successful = false;

        // LevenbergMarquardtStrategy::ComputeStep
        for (int j = 0; j < n; ++j) {
            const double s2 = scale[j] * scale[j];
            double d = s2 * hdiag[j];
            d = std::fmin(std::fmax(d, o.min_lm_diagonal), o.max_lm_diagonal);
            lambda[j] = d / (radius * s2);
        }
        bool ok = m.solve(lambda, delta);
        double model_cost_change = 0.0;
        if (ok) {
            double a = 0.0, b = 0.0;
            for (int j = 0; j < n; ++j) { a += delta[j] * lambda[j] * delta[j]; b += g[j] * delta[j]; if (!std::isfinite(delta[j])) ok = false; }
            model_cost_change = 0.5 * (a - b);
        }
        if (!ok || !(model_cost_change > 0.0)) {  // invalid step
            if (++invalid >= o.max_num_consecutive_invalid_steps) { out.termination = LM_FAILURE; break; }
            radius = radius / decrease_factor;  // StepIsInvalid == StepRejected
            decrease_factor *= 2.0;
            continue;
        }
        invalid = 0;

        const double cand = m.candidate_cost(delta);
        // ParameterToleranceReached
        if (m.step_norm() <= o.parameter_tolerance * (m.x_norm() + o.parameter_tolerance)) { out.termination = LM_CONVERGENCE; break; }
        // FunctionToleranceReached
        if (std::fabs(x_cost - cand) <= o.function_tolerance * x_cost) { out.termination = LM_CONVERGENCE; break; }

        const double relative_decrease = (x_cost - cand) / model_cost_change;
        if (relative_decrease > o.min_relative_decrease) {  // HandleSuccessfulStep
            m.accept();
            x_cost = m.linearize(g, hdiag);
            last_successful = true;
            ++out.num_successful_steps;
            const double t = 2.0 * relative_decrease - 1.0;
            radius = radius / std::fmax(1.0 / 3.0, 1.0 - t * t * t);
            radius = std::fmin(o.max_trust_region_radius, radius);
            decrease_factor = 2.0;
        } else {  // HandleUnsuccessfulStep
            radius = radius / decrease_factor;
            decrease_factor *= 2.0;
        }
    }
done:
    out.num_iterations = iter;
    out.final_cost = x_cost;
    out.final_radius = radius;
    out.total_time_s = elapsed();
}

// Dense Cholesky solve A x = b (A symmetric positive definite, row-major n x n, lower used).
// Destroys A.  Returns false if a pivot is not positive.
inline bool cholesky_solve(std::vector<double>& A, int n, std::vector<double>& b) {
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        A[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[(size_t)i * n + j];
            const double* ri = &A[(size_t)i * n];
            const double* rj = &A[(size_t)j * n];
            for (int k = 0; k < j; ++k) s -= ri[k] * rj[k];
            A[(size_t)i * n + j] = s / d;
        }
    }
    for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
    return true;
}

// Ceres HuberLoss(a) + Corrector [upstream]: s = |r|^2 ; rho = s (s<=a^2) else 2a sqrt(s) - a^2 ;
// rho'' <= 0 so the corrector is the pure sqrt(rho') scaling of r and J.  a <= 0: no loss.
inline void huber(double a, double s, double* rho, double* sqrt_rho1) {
    if (a <= 0.0) { *rho = s; *sqrt_rho1 = 1.0; return; }
    const double b = a * a;
    if (s > b) {
        const double r = std::sqrt(s);
        const double rho1 = std::fmax(std::numeric_limits<double>::min(), a / r);
        *rho = 2.0 * a * r - b;
        *sqrt_rho1 = std::sqrt(rho1);
    } else { *rho = s; *sqrt_rho1 = 1.0; }
}

}  // namespace oracle
