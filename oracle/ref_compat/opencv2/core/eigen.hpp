// oracle/ref_compat/opencv2/core/eigen.hpp -- TEST INFRASTRUCTURE ONLY: placeholder for the OpenCV header the reference's
// utility.h includes (nothing from it is used by the code compiled here).
#pragma once
