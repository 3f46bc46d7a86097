// oracle/ref_compat/ceres/ceres.h -- TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for <ceres/ceres.h> so that the reference's own factor headers
// (/root/reference/src/lvio_fusion/include/lvio_fusion/ceres/{base,visual_error,lidar_error,pose_error}.hpp) compile here
// without Ceres: the functors' operator() templates are then instantiated with the oracle's dual numbers (oracle/dual.h)
// by oracle/ref_harness.cpp.  Only the declarations those headers *name* are provided; AutoDiffCostFunction just keeps the
// functor (the harness differentiates by calling the functor with duals directly, which is what Ceres' autodiff does).
#pragma once
#include <cmath>
#include <utility>
#include <vector>
#include "../../dual.h"

namespace ceres {

using std::abs; using std::asin; using std::atan2; using std::cos; using std::sin; using std::sqrt;
// the oracle's dual-number math lives in namespace oracle and is found by argument-dependent lookup

class CostFunction {
public:
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const* const*, double*, double**) const { return false; }
    virtual int num_residuals() const { return 0; }
};
template <int kNumResiduals, int... Ns> class SizedCostFunction : public CostFunction {       // imu_error.hpp:12,124 derive from it
public:
    int num_residuals() const override { return kNumResiduals; }
};
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };
enum NumericDiffMethodType { CENTRAL, FORWARD, RIDDERS };
template <typename Functor, NumericDiffMethodType kMethod, int kNumResiduals, int... Ns>
class NumericDiffCostFunction : public CostFunction {                                            // named by imu_error.hpp:265 only
public:
    explicit NumericDiffCostFunction(Functor* f) : functor_(f) {}
    ~NumericDiffCostFunction() override { delete functor_; }
private:
    Functor* functor_;
};
template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
public:
    explicit AutoDiffCostFunction(Functor* f) : functor_(f) {}
    ~AutoDiffCostFunction() override { delete functor_; }
    const Functor& functor() const { return *functor_; }
    int num_residuals() const override { return kNumResiduals; }
    // residuals only (tests/cpp/ref_backend_dropin.cpp sums the cost of a recorded problem with the reference's functors)
    bool Evaluate(double const* const* p, double* r, double** j) const override { return j ? false : call(p, r, std::make_index_sequence<sizeof...(Ns)>()); }
private:
    template <size_t... I> bool call(double const* const* p, double* r, std::index_sequence<I...>) const { return (*functor_)(p[I]..., r); }
    Functor* functor_;
};


// ---- the problem-building surface src/association.cpp uses (adapt::Problem derives from ceres::Problem): recorded, not solved
class LossFunction { public: virtual ~LossFunction() {} virtual double huber_a() const { return 0.0; } };
class TrivialLoss : public LossFunction {};
class HuberLoss : public LossFunction { public: explicit HuberLoss(double a) : a_(a) {} double huber_a() const override { return a_; } private: double a_; };
class LocalParameterization { public: virtual ~LocalParameterization() {} };
class EigenQuaternionParameterization : public LocalParameterization {};                         // backend.cpp:99-101: named, recorded only
class IdentityParameterization : public LocalParameterization { public: explicit IdentityParameterization(int) {} };
class ProductParameterization : public LocalParameterization {
public:
    ProductParameterization(LocalParameterization* a, LocalParameterization* b) : a_(a), b_(b) {}
    ~ProductParameterization() override { delete a_; delete b_; }
private:
    LocalParameterization *a_, *b_;
};
struct ResidualBlock { CostFunction* cost; LossFunction* loss; std::vector<double*> blocks; };
typedef ResidualBlock* ResidualBlockId;
enum LinearSolverType { DENSE_QR, SPARSE_NORMAL_CHOLESKY, SPARSE_SCHUR };
class Solver { public: struct Options { LinearSolverType linear_solver_type; int max_num_iterations = 50; double max_solver_time_in_seconds = 1e9; int num_threads = 1; }; struct Summary { double final_cost = 0; int num_residual_blocks_reduced = 0; }; };
class Problem {
public:
    ~Problem() { for (ResidualBlock* r : residual_blocks) delete r; }
    void AddParameterBlock(double* v, int size) { parameter_blocks.push_back({v, size}); }
    void AddParameterBlock(double* v, int size, LocalParameterization*) { parameter_blocks.push_back({v, size}); }
    template <typename... Ts> ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, Ts*... xs) {
        residual_blocks.push_back(new ResidualBlock{cost, loss, {x0, xs...}});
        return residual_blocks.back();
    }
    void SetParameterBlockConstant(double* v) { constant_blocks.push_back(v); }          // environment.cpp:62-68: recorded only
    void GetResidualBlocksForParameterBlock(const double* v, std::vector<ResidualBlockId>* out) const {
        out->clear();
        for (ResidualBlock* r : residual_blocks) for (double* b : r->blocks) if (b == v) { out->push_back(r); break; }
    }
    std::vector<ResidualBlock*> residual_blocks;
    std::vector<double*> constant_blocks;
    std::vector<std::pair<double*, int>> parameter_blocks;
};
void Solve(const Solver::Options&, Problem*, Solver::Summary*);      // named by adapt::Solve, never called here

}  // namespace ceres
