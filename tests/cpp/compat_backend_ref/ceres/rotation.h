#pragma once
#include "../../../../oracle/ref_compat/ceres/rotation.h"
