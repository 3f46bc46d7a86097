// oracle/ref_compat/lvio_fusion/frame.h -- TEST INFRASTRUCTURE ONLY: src/preintegration.cpp includes the reference's
// frame.h without using anything from it; the real header drags in the whole frontend (OpenCV, PCL, ...).
#pragma once
