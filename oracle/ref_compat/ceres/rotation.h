// oracle/ref_compat/ceres/rotation.h -- TEST INFRASTRUCTURE ONLY.
// The three <ceres/rotation.h> templates the reference's base.hpp calls (base.hpp:30,63,82), w-first quaternions.
// [upstream, Ceres 2.x, not in the reference tree]: QuaternionRotatePoint normalises q, UnitQuaternionRotatePoint uses
// the "uv" form, QuaternionProduct is the Hamilton product.
#pragma once
#include "ceres.h"

namespace ceres {

template <typename T> inline T DotProduct(const T x[3], const T y[3]) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }
template <typename T> inline void QuaternionProduct(const T z[4], const T w[4], T zw[4]) {
    zw[0] = z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3];
    zw[1] = z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2];
    zw[2] = z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1];
    zw[3] = z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0];
}
template <typename T> inline void UnitQuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
    T uv0 = q[2] * pt[2] - q[3] * pt[1], uv1 = q[3] * pt[0] - q[1] * pt[2], uv2 = q[1] * pt[1] - q[2] * pt[0];
    uv0 += uv0; uv1 += uv1; uv2 += uv2;
    result[0] = pt[0] + q[0] * uv0; result[1] = pt[1] + q[0] * uv1; result[2] = pt[2] + q[0] * uv2;
    result[0] += q[2] * uv2 - q[3] * uv1; result[1] += q[3] * uv0 - q[1] * uv2; result[2] += q[1] * uv1 - q[2] * uv0;
}
template <typename T> inline void QuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
    const T scale = T(1) / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const T unit[4] = {scale * q[0], scale * q[1], scale * q[2], scale * q[3]};
    UnitQuaternionRotatePoint(unit, pt, result);
}

}  // namespace ceres
