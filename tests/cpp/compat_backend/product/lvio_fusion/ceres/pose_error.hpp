// the reference's factor header is replaced by the product's device factor records.
#pragma once
#include "lvio_b200/factors.h"
#include "lvio_fusion/common.h"
