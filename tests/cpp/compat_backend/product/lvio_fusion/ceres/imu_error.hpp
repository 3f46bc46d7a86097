// the reference's factor header is replaced by the product's device factor records (the
// reference's header also brings imu/preintegration.h and utility.h into scope for its includers).
#pragma once
#include "lvio_b200/factors.h"
#include "lvio_fusion/common.h"
#include "lvio_fusion/imu/preintegration.h"
#include "lvio_fusion/utility.h"
