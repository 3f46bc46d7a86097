// oracle/ref_compat/lvio_fusion/map.h -- TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's Map singleton: the two
// members FeatureAssociation::AddScan / UndistortPoint name (never called by the harness).
#pragma once
#include "lvio_fusion/frame.h"

namespace lvio_fusion {
class Map {
public:
    static Map& Instance() { static Map m; return m; }
    Frames GetKeyFrames(double, double = 0, int = 0) { return keyframes; }
    SE3d ComputePose(double) { return SE3d(); }
    Frames keyframes;
};
}  // namespace lvio_fusion
