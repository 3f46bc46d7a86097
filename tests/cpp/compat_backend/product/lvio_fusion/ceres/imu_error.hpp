// the reference's factor header is replaced by the product's device factor records (the
// reference's header also brings imu/preintegration.h and utility.h into scope for its includers).
#pragma once
#include "lvio_b200/factors.h"
#include "lvio_fusion/common.h"
#include "lvio_fusion/imu/preintegration.h"
#include "lvio_fusion/utility.h"
namespace lvio_fusion {
// imu::InertialOptimization (src/tools.cpp:35-90, the gravity-direction step of IMU initialisation) is outside the drop-in:
// its NumericDiff factor is only named here so that the translation unit compiles; the harness never reaches it.
class ImuInitGError {
public:
    template <class... A> static ceres::CostFunction* Create(A&&...) { std::abort(); return nullptr; }
};
}  // namespace lvio_fusion
