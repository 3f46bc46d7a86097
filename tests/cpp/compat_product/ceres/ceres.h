// tests/cpp/compat_product/ceres/ceres.h -- lets the reference's own headers (`#include <ceres/ceres.h>`) resolve to the
// PRODUCT shim, so that tests/cpp/ref_navsat_dropin.cpp can compile the reference's navsat functors against it unchanged.
#pragma once
#include "lvio_b200/ceres_shim.h"
#include "lvio_b200/ceres_autodiff.h"
