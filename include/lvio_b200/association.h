// include/lvio_b200/association.h -- scan-to-map registration with the call shape of
// lvio_fusion::FeatureAssociation::ScanToMapWithGround / ScanToMapWithSegmented
// (/root/reference/src/lvio_fusion/include/lvio_fusion/lidar/association.h:30-32, src/association.cpp:270-384).
//
// In the reference each call builds a pcl::KdTreeFLANN on the map cloud, searches 3 neighbours per scan point on
// the host and adds one LidarPlaneErrorRPZ/YXY residual block per accepted point; the caller then runs
// adapt::Solve (mapping.cpp:159-177).  Here the call uploads the map cloud once per call into the voxel hash,
// and adds ONE lvb::ScanToMapCost describing the whole association; ceres::Solve (ceres_shim.h) runs
// association + LM on the device and updates para[] in place.  Frame types are duck-typed:
//   frame->pose.data()                       Sophus::SE3d layout [qx qy qz qw tx ty tz]
//   frame->feature_lidar->points_ground / points_surf : contiguous clouds of 32-byte pcl::PointXYZI
//       (anything with .size() and .data() whose records start with float x, y, z)
//   frame->weights.lidar_ground / lidar_surf / visual,  frame->features_left.size()
#pragma once
#include <cmath>
#include <type_traits>
#include <unordered_map>
#include "ceres_shim.h"

// Inside the reference tree the class of the same name already exists (it also owns the feature extraction):
// compile with -DLVB_ASSOCIATION_CLASS=ScanToMapDevice and forward the two members to it.
#ifndef LVB_ASSOCIATION_CLASS
#define LVB_ASSOCIATION_CLASS FeatureAssociation
#endif

namespace lvio_fusion {

namespace assoc_detail {
// first record of a cloud: pcl::PointCloud keeps its records in the public vector `points` (every PCL version; PCL >= 1.11
// also has data()); plain containers have data()
template <class Cloud> inline auto cloud_data(const Cloud& c, int) -> decltype(c.points.data()) { return c.points.data(); }
template <class Cloud> inline auto cloud_data(const Cloud& c, long) -> decltype(c.data()) { return c.data(); }
// adapt::Problem (adapt/problem.h:37-47) hides ceres::Problem::AddResidualBlock behind an overload that takes the block's
// ProblemType first and keeps a census of them; a plain ceres::Problem has no such census.  Use whichever the caller has:
// `lidar` selects ProblemType::LidarError, otherwise ProblemType::Other (association.cpp:318,324).
template <class P>
inline auto add_block(P& p, int, bool lidar, ceres::CostFunction* c, ceres::LossFunction* l, double* x0, double* x1, double* x2) -> decltype(p.num_types, void()) {
    typedef typename std::decay<decltype(p.num_types.begin()->first)>::type Type;
    p.AddResidualBlock(lidar ? Type::LidarError : Type::Other, c, l, x0, x1, x2);
}
template <class P>
inline void add_block(P& p, long, bool, ceres::CostFunction* c, ceres::LossFunction* l, double* x0, double* x1, double* x2) { p.AddResidualBlock(c, l, x0, x1, x2); }
}  // namespace assoc_detail

class LVB_ASSOCIATION_CLASS {
public:
    explicit LVB_ASSOCIATION_CLASS(double lidar_resolution = 0.2) : resolution_(lidar_resolution) {}
    ~LVB_ASSOCIATION_CLASS() {                      // the calling thread's handles; another thread's go with that thread's context
        auto& m = handles_of_thread();
        auto it = m.find(this);
        if (it != m.end()) { if (it->second.ground) lvb_icp_destroy(it->second.ground); if (it->second.surf) lvb_icp_destroy(it->second.surf); m.erase(it); }
    }

    // association.cpp:270-326 : pitch/roll/z = para+1,+2,+5 ; gate d2 < 100 res^2 ; TrivialLoss ; PoseErrorRPZ prior
    template <class FramePtr, class Problem>
    bool ScanToMapWithGround(FramePtr frame, FramePtr map_frame, double* para, Problem& problem, bool relocate = false) {
        return add(0, frame, map_frame, frame->feature_lidar->points_ground, map_frame->feature_lidar->points_ground, para, problem, relocate,
                   resolution_ * resolution_ * 100, frame->weights.lidar_ground, new ceres::TrivialLoss(), handles_of_thread()[this].ground);
    }
    // association.cpp:328-384 : yaw/x/y = para+0,+3,+4 ; gate d2 < 25 res^2 ; HuberLoss(0.1) ; PoseErrorYXY prior
    template <class FramePtr, class Problem>
    bool ScanToMapWithSegmented(FramePtr frame, FramePtr map_frame, double* para, Problem& problem, bool relocate = false) {
        return add(1, frame, map_frame, frame->feature_lidar->points_surf, map_frame->feature_lidar->points_surf, para, problem, relocate,
                   resolution_ * resolution_ * 25, frame->weights.lidar_surf, new ceres::HuberLoss(0.1), handles_of_thread()[this].surf);
    }

    // ---- device-resident map.  Mapping::ToWorld (mapping.cpp:205-220) -> AppendKeyframe: the keyframe's two
    // robot-frame feature clouds go to the device once and stay there in the world frame; Mapping::BuildMapFrame (:114-137) ->
    // BuildMapFrame: merge of the given keyframes in order, SegmentGround on the merged ground cloud, voxel hash -- all on the
    // device.  After a BuildMapFrame the two ScanToMapWith* members register against that resident map frame and no longer upload
    // map_frame's clouds; UsePerCallMap() goes back to the per-call behaviour (Relocator::RelocateByPoints hands over its own map).
    template <class Cloud>
    bool AppendKeyframe(long long key, const Cloud& ground_robot, const Cloud& surf_robot, const double* pose7) {
        Handles& h = handles_of_thread()[this];
        if (!ensure_handles(h)) return false;
        const int stride = (int)sizeof(assoc_detail::cloud_data(ground_robot, 0)[0]);
        lvb_icp_map_evict(h.ground, key); lvb_icp_map_evict(h.surf, key);            // re-registering a keyframe (ToWorld(start)) replaces its clouds
        return lvb_icp_map_append(h.ground, key, assoc_detail::cloud_data(ground_robot, 0), (int)ground_robot.size(), stride, pose7) == LVB_OK &&
               lvb_icp_map_append(h.surf, key, assoc_detail::cloud_data(surf_robot, 0), (int)surf_robot.size(), stride, pose7) == LVB_OK;
    }
    bool EvictKeyframe(long long key) {
        Handles& h = handles_of_thread()[this];
        return ensure_handles(h) && lvb_icp_map_evict(h.ground, key) == LVB_OK && lvb_icp_map_evict(h.surf, key) == LVB_OK;
    }
    bool BuildMapFrame(const long long* keys, int n_keys, double ground_ransac_threshold) {
        Handles& h = handles_of_thread()[this];
        if (!ensure_handles(h)) return false;
        h.resident = lvb_icp_map_build(h.ground, keys, n_keys, cell_of(resolution_ * resolution_ * 100), ground_ransac_threshold, nullptr) == LVB_OK &&
                     lvb_icp_map_build(h.surf, keys, n_keys, cell_of(resolution_ * resolution_ * 25), -1.0, nullptr) == LVB_OK;
        return h.resident;
    }
    void UsePerCallMap() { handles_of_thread()[this].resident = false; }

private:
    static float cell_of(double thr) { return std::nextafter((float)std::sqrt(thr), 1e30f) * 1.0001f; }
    template <class FramePtr, class Cloud, class Problem>
    bool add(int mode, FramePtr frame, FramePtr map_frame, const Cloud& scan, const Cloud& map, double* para, Problem& problem, bool relocate,
             double thr, double weight, ceres::LossFunction* loss, lvb_icp*& icp) {
        lvb::Runtime& rt = lvb::Runtime::get();
        if (!rt.ensure()) { delete loss; return false; }
        if (!icp && lvb_icp_create(rt.ctx, &icp) != LVB_OK) { delete loss; return false; }
        const int stride = (int)sizeof(assoc_detail::cloud_data(scan, 0)[0]);
        if (!handles_of_thread()[this].resident &&
            lvb_icp_set_map(icp, assoc_detail::cloud_data(map, 0), (int)map.size(), stride, cell_of(thr)) != LVB_OK) { delete loss; return false; }
        if (mode == 0) { problem.AddParameterBlock(para + 1, 1); problem.AddParameterBlock(para + 2, 1); problem.AddParameterBlock(para + 5, 1); }
        else { problem.AddParameterBlock(para + 0, 1); problem.AddParameterBlock(para + 3, 1); problem.AddParameterBlock(para + 4, 1); }
        double* x0 = mode == 0 ? para + 1 : para + 0; double* x1 = mode == 0 ? para + 2 : para + 3; double* x2 = mode == 0 ? para + 5 : para + 4;
        lvb::ScanToMapCost* c = new lvb::ScanToMapCost();
        c->mode = mode; c->icp = icp; c->scan = assoc_detail::cloud_data(scan, 0); c->n = (int)scan.size(); c->stride = stride;
        std::memcpy(c->frame_pose, frame->pose.data(), sizeof(c->frame_pose)); std::memcpy(c->map_pose, map_frame->pose.data(), sizeof(c->map_pose));
        c->rpyxyz = para; c->weight = weight; c->dist_thr = thr;
        assoc_detail::add_block(problem, 0, true, c, loss, x0, x1, x2);
        if (!relocate) {
            lvb::IcpPriorCost* p = new lvb::IcpPriorCost();
            p->mode = mode; p->weight = (double)frame->features_left.size() * frame->weights.visual;   // association.cpp:323,381
            assoc_detail::add_block(problem, 0, false, p, nullptr, x0, x1, x2);
        }
        return true;
    }
    // lvb_icp handles (voxel hash + scratch on one context's stream) are single-threaded and the context is per host thread
    // (ceres_shim.h): Mapping::Optimize and Relocator::RelocateByPoints reach the same object from two threads without a lock
    // (mapping.cpp:155-177, relocator.cpp:188-206), so every thread gets its own pair.
    struct Handles { lvb_icp* ground = nullptr; lvb_icp* surf = nullptr; bool resident = false; };
    static bool ensure_handles(Handles& h) {
        lvb::Runtime& rt = lvb::Runtime::get();
        if (!rt.ensure()) return false;
        return (h.ground || lvb_icp_create(rt.ctx, &h.ground) == LVB_OK) && (h.surf || lvb_icp_create(rt.ctx, &h.surf) == LVB_OK);
    }
    static std::unordered_map<const void*, Handles>& handles_of_thread() { static thread_local std::unordered_map<const void*, Handles> m; return m; }
    double resolution_;
};

}  // namespace lvio_fusion
