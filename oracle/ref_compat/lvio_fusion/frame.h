// oracle/ref_compat/lvio_fusion/frame.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/lvio_fusion/frame.h (which drags in the whole frontend): the members
// src/association.cpp and src/preintegration.cpp touch.
#pragma once
#include "lvio_fusion/common.h"
#include "lvio_fusion/lidar/feature.h"

namespace lvio_fusion {
struct Weights { double visual = 1, lidar_ground = 1, lidar_surf = 1; };
class Frame {
public:
    typedef std::shared_ptr<Frame> Ptr;
    static Frame::Ptr Create() { return Frame::Ptr(new Frame); }
    double time = 0;
    std::map<unsigned long, int> features_left;          // only its size() is read (association.cpp:323,381)
    lidar::Feature::Ptr feature_lidar;
    Weights weights;
    SE3d pose;
    Vector3d t() { return pose.translation(); }
};
typedef std::map<double, Frame::Ptr> Frames;
}  // namespace lvio_fusion
