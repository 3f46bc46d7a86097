// oracle/ref_assoc_harness.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the REFERENCE's own FeatureAssociation code (src/association.cpp + src/projection.cpp, compiled in place):
//   extract : Preprocess -> ImageProjection::Process -> AdjustDistortion -> CalculateSmoothness -> ExtractFeatures
//             (association.cpp:88-268) on a raw sweep; dumps the segmented cloud (with the relative time in the intensity),
//             the curvatures and the two feature clouds;
//   scan2map: ScanToMapWithGround / ScanToMapWithSegmented (association.cpp:270-384) on given scan / map feature clouds and
//             poses; dumps, per scan point, whether a LidarPlaneError block was created and its residual / 3 Jacobian
//             entries at `para` (the functor instantiated with duals, as AutoDiffCostFunction does), plus the prior weight.
// PCL's filters / kd-tree are the adapters of oracle/ref_compat/pcl_standin.h over the oracle's restatements: what this pins
// is the reference's own loops, gates, weights and factor creation.
#include <cstdio>
#include <cstring>
#include <vector>

#define private public          // the harness calls the pipeline stages one by one (same class layout in both translation units)
#include "lvio_fusion/lidar/association.h"
#include "lvio_fusion/ceres/lidar_error.hpp"
#include "lvio_fusion/ceres/pose_error.hpp"
#undef private
#include "lvio_fusion/lidar/lidar.h"

namespace lvio_fusion { std::vector<Lidar::Ptr> Lidar::devices_; }
const double epsilon = 1e-3;
const int num_threads = 1;
using namespace lvio_fusion;
using oracle::Dual;

static void read_all(const char* path, std::vector<unsigned char>& buf) {
    FILE* f = fopen(path, "rb"); if (!f) exit(2);
    fseek(f, 0, SEEK_END); buf.resize((size_t)ftell(f)); fseek(f, 0, SEEK_SET);
    if (fread(buf.data(), 1, buf.size(), f) != buf.size()) exit(2);
    fclose(f);
}
static void put_cloud(FILE* o, const PointICloud& c) {
    const int m = (int)c.size(); fwrite(&m, sizeof(int), 1, o);
    for (int i = 0; i < m; ++i) { const float v[4] = {c[i].x, c[i].y, c[i].z, c[i].intensity}; fwrite(v, sizeof(float), 4, o); }
}
static PointICloud get_cloud(const unsigned char*& p) {
    int m; memcpy(&m, p, 4); p += 4;
    PointICloud c;
    for (int i = 0; i < m; ++i) { float v[4]; memcpy(v, p, 16); p += 16; PointI q; q.x = v[0]; q.y = v[1]; q.z = v[2]; q.intensity = v[3]; c.push_back(q); }
    return c;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: ref_assoc extract|scan2map <in.bin> <out.bin>\n"); return 64; }
    std::vector<unsigned char> buf; read_all(argv[2], buf);
    const unsigned char* p = buf.data();
    double cfg[10];      // num_scans horizon_scan ang_res_y ang_bottom ground_rows cycle_time min_range max_range resolution (+1 spare)
    memcpy(cfg, p, sizeof(cfg)); p += sizeof(cfg);
    double ext[7]; memcpy(ext, p, sizeof(ext)); p += sizeof(ext);
    Lidar::Create(cfg[8], SE3d(ext));
    FeatureAssociation fa((int)cfg[0], (int)cfg[1], cfg[2], cfg[3], (int)cfg[4], cfg[5], cfg[6], cfg[7], 0, 0);
    FILE* o = fopen(argv[3], "wb"); if (!o) return 2;
    if (std::string(argv[1]) == "extract") {
        int n; memcpy(&n, p, 4); p += 4;
        PointICloud points;
        for (int i = 0; i < n; ++i) { float v[3]; memcpy(v, p, 12); p += 12; PointI q; q.x = v[0]; q.y = v[1]; q.z = v[2]; points.push_back(q); }
        Frame::Ptr frame = Frame::Create();
        fa.Preprocess(points);
        PointICloud seg;
        SegmentedInfo info = fa.projection_->Process(points, seg);
        fa.AdjustDistortion(seg, info);
        fa.CalculateSmoothness(seg, info);
        put_cloud(o, seg);
        fwrite(fa.curvatures, sizeof(float), seg.size(), o);          // entries outside [5, size-5) are whatever the array held
        fa.ExtractFeatures(seg, info, frame);
        put_cloud(o, frame->feature_lidar->points_ground);
        put_cloud(o, frame->feature_lidar->points_surf);
    } else {
        // mode, frame pose, map pose, rpyxyz, weights (visual, ground, surf), n_features_left, relocate, scan cloud, map cloud
        int mode; memcpy(&mode, p, 4); p += 4;
        double fp[7], mp[7], para[6], w[3]; memcpy(fp, p, 56); p += 56; memcpy(mp, p, 56); p += 56; memcpy(para, p, 48); p += 48; memcpy(w, p, 24); p += 24;
        int nfeat, relocate; memcpy(&nfeat, p, 4); p += 4; memcpy(&relocate, p, 4); p += 4;
        Frame::Ptr frame = Frame::Create(), map_frame = Frame::Create();
        frame->pose = SE3d(fp); map_frame->pose = SE3d(mp);
        frame->weights.visual = w[0]; frame->weights.lidar_ground = w[1]; frame->weights.lidar_surf = w[2];
        for (int i = 0; i < nfeat; ++i) frame->features_left[(unsigned long)i] = 0;
        frame->feature_lidar = lidar::Feature::Create(); map_frame->feature_lidar = lidar::Feature::Create();
        const PointICloud scan = get_cloud(p), map = get_cloud(p);
        (mode == 0 ? frame->feature_lidar->points_ground : frame->feature_lidar->points_surf) = scan;
        (mode == 0 ? map_frame->feature_lidar->points_ground : map_frame->feature_lidar->points_surf) = map;
        adapt::Problem problem;
        if (mode == 0) fa.ScanToMapWithGround(frame, map_frame, para, problem, relocate != 0);
        else fa.ScanToMapWithSegmented(frame, map_frame, para, problem, relocate != 0);
        // the residual blocks were added in scan order; match them back to scan points through the stored point p_
        const int n = (int)scan.size();
        std::vector<double> out((size_t)n * 5, 0.0);          // accepted, r, J[3]
        size_t next_block = 0;
        double prior_w = -1.0, loss_a = 0.0;
        std::vector<ceres::ResidualBlock*>& rb = problem.residual_blocks;
        for (int i = 0; i < n && next_block < rb.size(); ++i) {
            const double x[3] = {*rb[next_block]->blocks[0], *rb[next_block]->blocks[1], *rb[next_block]->blocks[2]};
            Dual<3> X[3] = {Dual<3>::seed(x[0], 0), Dual<3>::seed(x[1], 1), Dual<3>::seed(x[2], 2)}, Y[1];
            bool mine = false;
            if (mode == 0) {
                auto* c = dynamic_cast<ceres::AutoDiffCostFunction<LidarPlaneErrorRPZ, 1, 1, 1, 1>*>(rb[next_block]->cost);
                if (c && c->functor().origin_error_.p_.x() == (double)scan[i].x && c->functor().origin_error_.p_.y() == (double)scan[i].y && c->functor().origin_error_.p_.z() == (double)scan[i].z) { c->functor()(X, X + 1, X + 2, Y); mine = true; }
            } else {
                auto* c = dynamic_cast<ceres::AutoDiffCostFunction<LidarPlaneErrorYXY, 1, 1, 1, 1>*>(rb[next_block]->cost);
                if (c && c->functor().origin_error_.p_.x() == (double)scan[i].x && c->functor().origin_error_.p_.y() == (double)scan[i].y && c->functor().origin_error_.p_.z() == (double)scan[i].z) { c->functor()(X, X + 1, X + 2, Y); mine = true; }
            }
            if (!mine) continue;
            out[(size_t)i * 5] = 1.0; out[(size_t)i * 5 + 1] = Y[0].v;
            for (int k = 0; k < 3; ++k) out[(size_t)i * 5 + 2 + k] = Y[0].d[k];
            loss_a = rb[next_block]->loss ? rb[next_block]->loss->huber_a() : 0.0;
            ++next_block;
        }
        int n_lidar_blocks = (int)next_block, n_other = (int)(rb.size() - next_block);
        if (n_other == 1) {          // the PoseErrorRPZ / YXY prior: its weight is the residual slope
            Dual<3> X[3] = {Dual<3>::seed(*rb.back()->blocks[0], 0), Dual<3>::seed(*rb.back()->blocks[1], 1), Dual<3>::seed(*rb.back()->blocks[2], 2)}, Y[3];
            if (mode == 0) dynamic_cast<ceres::AutoDiffCostFunction<PoseErrorRPZ, 3, 1, 1, 1>*>(rb.back()->cost)->functor()(X, X + 1, X + 2, Y);
            else dynamic_cast<ceres::AutoDiffCostFunction<PoseErrorYXY, 3, 1, 1, 1>*>(rb.back()->cost)->functor()(X, X + 1, X + 2, Y);
            prior_w = Y[2].d[2];
        }
        const double head[4] = {(double)n_lidar_blocks, (double)n_other, prior_w, loss_a};
        fwrite(head, sizeof(double), 4, o);
        fwrite(out.data(), sizeof(double), out.size(), o);
    }
    fclose(o);
    return 0;
}
