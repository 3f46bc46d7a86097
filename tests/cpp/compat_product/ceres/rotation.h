// see ceres.h in this directory
#pragma once
#include "lvio_b200/ceres_autodiff.h"
