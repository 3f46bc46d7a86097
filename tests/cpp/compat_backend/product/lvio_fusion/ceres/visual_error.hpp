// the reference's factor header is replaced by the product's device factor records (the
// reference's header also brings visual/camera.h into scope for its includers).
#pragma once
#include "lvio_b200/factors.h"
#include "lvio_fusion/visual/camera.h"
