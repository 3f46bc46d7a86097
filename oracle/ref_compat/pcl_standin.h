// oracle/ref_compat/pcl_standin.h -- TEST INFRASTRUCTURE ONLY.
// The PCL classes src/association.cpp instantiates, as thin adapters over the oracle's own restatements of those
// algorithms (oracle/lidar.h: VoxelGrid, RadiusOutlierRemoval, plane RANSAC; oracle/icp.h: exact 3-NN).  Compiling the
// reference against them pins the reference's OWN loops (relative time, smoothness, feature split, association gate, factor
// creation); PCL's internals stay "[upstream], unpinned".
#pragma once
#include "lvio_fusion/common.h"
#include "../lidar.h"

namespace pcl {

template <class PIn, class POut> inline void copyPointCloud(const PointCloud<PIn>& in, PointCloud<POut>& out) {
    PointCloud<POut> tmp;
    tmp.header = in.header;
    for (const PIn& p : in.points) { POut q; q.x = p.x; q.y = p.y; q.z = p.z; copy_intensity(p, q); tmp.points.push_back(q); }
    tmp.width = (unsigned int)tmp.points.size(); tmp.height = 1; tmp.is_dense = in.is_dense;
    out = tmp;
}
template <class PointT> inline void removeNaNFromPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, std::vector<int>& index) {
    PointCloud<PointT> tmp; index.clear();
    for (size_t i = 0; i < in.points.size(); ++i) { const PointT& p = in.points[i]; if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) { tmp.points.push_back(p); index.push_back((int)i); } }
    tmp.header = in.header; tmp.width = (unsigned int)tmp.points.size(); tmp.height = 1; tmp.is_dense = true;
    out = tmp;
}

inline std::vector<oracle::PointI> to_oracle(const PointCloud<PointXYZI>& c) { std::vector<oracle::PointI> v; for (const PointXYZI& p : c.points) v.push_back(oracle::PointI{p.x, p.y, p.z, p.intensity}); return v; }
inline void from_oracle(const std::vector<oracle::PointI>& v, PointCloud<PointXYZI>& c) {
    c.points.clear();
    for (const oracle::PointI& p : v) { PointXYZI q; q.x = p.x; q.y = p.y; q.z = p.z; q.intensity = p.intensity; c.points.push_back(q); }
    c.width = (unsigned int)c.points.size(); c.height = 1;
}

template <class PointT> class VoxelGrid {
public:
    void setLeafSize(float lx, float, float) { leaf_ = lx; }
    void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { in_ = c; }
    void filter(PointCloud<PointT>& out) { from_oracle(oracle::voxel_grid(to_oracle(*in_), leaf_), out); }
private:
    float leaf_ = 1; typename PointCloud<PointT>::Ptr in_;
};
// Mapping::GetGlobalMap (src/mapping.cpp:228-244) down-samples the coloured display cloud: visualisation, outside the path
template <> class VoxelGrid<PointXYZRGB> {
public:
    void setLeafSize(float, float, float) {}
    void setInputCloud(const PointCloud<PointXYZRGB>::Ptr& c) { in_ = c; }
    void filter(PointCloud<PointXYZRGB>& out) { out = *in_; }
private:
    PointCloud<PointXYZRGB>::Ptr in_;
};
template <class PointT> class RadiusOutlierRemoval {
public:
    void setRadiusSearch(double r) { r_ = r; }
    void setMinNeighborsInRadius(int k) { k_ = k; }
    void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { in_ = c; }
    void filter(PointCloud<PointT>& out) { from_oracle(oracle::radius_outlier_removal(to_oracle(*in_), r_, k_), out); }
private:
    double r_ = 0; int k_ = 1; typename PointCloud<PointT>::Ptr in_;
};
struct ModelCoefficients { std::vector<float> values; };
struct PointIndices { typedef std::shared_ptr<PointIndices> Ptr; std::vector<int> indices; };
enum { SACMODEL_PLANE = 0 };
enum { SAC_RANSAC = 0 };
template <class PointT> class SACSegmentation {
public:
    void setOptimizeCoefficients(bool) {}
    void setModelType(int) {}
    void setMethodType(int) {}
    void setMaxIterations(int) {}
    void setDistanceThreshold(double t) { thr_ = t; }
    void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { in_ = c; }
    void segment(PointIndices& inliers, ModelCoefficients&) {       // indices of the oracle's RANSAC inliers (input order)
        const std::vector<oracle::PointI> v = to_oracle(*in_), keep = oracle::segment_ground(v, thr_);
        inliers.indices.clear();
        size_t k = 0;
        for (size_t i = 0; i < v.size() && k < keep.size(); ++i) if (v[i].x == keep[k].x && v[i].y == keep[k].y && v[i].z == keep[k].z && v[i].intensity == keep[k].intensity) { inliers.indices.push_back((int)i); ++k; }
    }
private:
    double thr_ = 0; typename PointCloud<PointT>::Ptr in_;
};
template <class PointT> class ExtractIndices {
public:
    void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { in_ = c; }
    void setIndices(const PointIndices::Ptr& i) { idx_ = i; }
    void setNegative(bool n) { neg_ = n; }
    void filter(PointCloud<PointT>& out) {
        PointCloud<PointT> tmp;
        std::vector<char> sel(in_->points.size(), 0);
        for (int i : idx_->indices) sel[(size_t)i] = 1;
        for (size_t i = 0; i < in_->points.size(); ++i) if ((sel[i] != 0) != neg_) tmp.points.push_back(in_->points[i]);
        tmp.width = (unsigned int)tmp.points.size(); tmp.height = 1;
        out = tmp;
    }
private:
    typename PointCloud<PointT>::Ptr in_; PointIndices::Ptr idx_; bool neg_ = false;
};
// exact k-NN, float32 squared distances without FMA, ascending (d2, index): the oracle's definition (oracle/icp.h)
template <class PointT> class KdTreeFLANN {
public:
    void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { pts_.clear(); for (const PointT& p : c->points) pts_.push_back(oracle::Pt{p.x, p.y, p.z}); }
    int nearestKSearch(const PointT& q, int k, std::vector<int>& idx, std::vector<float>& d2) const {
        idx.assign((size_t)k, 0); d2.assign((size_t)k, 0.0f);
        const oracle::Knn3 r = oracle::knn3_brute(pts_.data(), (int)pts_.size(), oracle::Pt{q.x, q.y, q.z});
        for (int i = 0; i < k && i < 3; ++i) { idx[(size_t)i] = r.idx[i]; d2[(size_t)i] = r.d2[i]; }
        return k;
    }
private:
    std::vector<oracle::Pt> pts_;
};

}  // namespace pcl
