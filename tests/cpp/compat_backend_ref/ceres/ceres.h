// tests/cpp/compat_backend_ref/ceres/ceres.h -- TEST INFRASTRUCTURE ONLY.  For the "truth" build of
// tests/cpp/ref_backend_dropin.cpp: <ceres/ceres.h> is the recording stand-in of oracle/ref_compat, so that the reference's
// OWN factor headers are the ones src/backend.cpp and src/tools.cpp instantiate.
#pragma once
#include "../../../../oracle/ref_compat/ceres/ceres.h"
