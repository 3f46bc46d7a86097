// oracle/ref_compat/ceres/ceres.h -- TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for <ceres/ceres.h> so that the reference's own factor headers
// (/root/reference/src/lvio_fusion/include/lvio_fusion/ceres/{base,visual_error,lidar_error,pose_error}.hpp) compile here
// without Ceres: the functors' operator() templates are then instantiated with the oracle's dual numbers (oracle/dual.h)
// by oracle/ref_harness.cpp.  Only the declarations those headers *name* are provided; AutoDiffCostFunction just keeps the
// functor (the harness differentiates by calling the functor with duals directly, which is what Ceres' autodiff does).
#pragma once
#include <cmath>
#include "../../dual.h"

namespace ceres {

using std::abs; using std::asin; using std::atan2; using std::cos; using std::sin; using std::sqrt;
// the oracle's dual-number math lives in namespace oracle and is found by argument-dependent lookup

class CostFunction {
public:
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const* const*, double*, double**) const { return false; }
};
template <int kNumResiduals, int... Ns> class SizedCostFunction : public CostFunction {};      // imu_error.hpp:12,124 derive from it
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
private:
    Functor* functor_;
};

}  // namespace ceres
