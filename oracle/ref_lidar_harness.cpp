// oracle/ref_lidar_harness.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the REFERENCE's own range-image code on a raw sweep: filter_points_by_distance (include/lvio_fusion/utility.h:70-96)
// and ImageProjection::Process (src/projection.cpp, compiled in place with the stand-ins of oracle/ref_compat for cv::Mat
// and pcl::PointCloud).  Whatever the toolchain decides for the unqualified abs / atan2 / sqrt calls on floats in that file
// is what gets pinned.  Output: segmented cloud (x y z intensity), range, ground flag, column, ring start / end, orientation.
#include <cmath>
#include <cstdio>
#include <vector>

#include "lvio_fusion/lidar/projection.h"
#include "lvio_fusion/utility.h"

using namespace lvio_fusion;

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_lidar <in.bin> <out.bin>\n"); return 64; }
    FILE* f = fopen(argv[1], "rb"); if (!f) return 2;
    double cfg[7];      // num_scans horizon_scan ang_res_y ang_bottom ground_rows min_range max_range
    int n = 0;
    if (fread(cfg, sizeof(double), 7, f) != 7 || fread(&n, sizeof(int), 1, f) != 1) return 2;
    std::vector<float> xyz(3 * (size_t)n);
    if (fread(xyz.data(), sizeof(float), xyz.size(), f) != xyz.size()) return 2;
    fclose(f);
    PointICloud points;
    for (int i = 0; i < n; ++i) {        // pcl::removeNaNFromPointCloud (association.cpp:98-99)
        if (!std::isfinite(xyz[3 * i]) || !std::isfinite(xyz[3 * i + 1]) || !std::isfinite(xyz[3 * i + 2])) continue;
        PointI p; p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
        points.push_back(p);
    }
    filter_points_by_distance(points, points, (float)cfg[5], (float)cfg[6]);      // min_range_, max_range_ are doubles narrowed at the call
    const int R = (int)cfg[0], W = (int)cfg[1];
    ImageProjection proj(R, W, cfg[2], cfg[3], (int)cfg[4]);
    PointICloud seg;
    SegmentedInfo info = proj.Process(points, seg);
    FILE* o = fopen(argv[2], "wb"); if (!o) return 2;
    const int m = (int)seg.size();
    fwrite(&m, sizeof(int), 1, o);
    for (int i = 0; i < m; ++i) { const float v[4] = {seg[i].x, seg[i].y, seg[i].z, seg[i].intensity}; fwrite(v, sizeof(float), 4, o); }
    for (int i = 0; i < m; ++i) { const float r = info.range[i]; fwrite(&r, sizeof(float), 1, o); }
    for (int i = 0; i < m; ++i) { const unsigned char g = info.ground_flag[i] ? 1 : 0; fwrite(&g, 1, 1, o); }
    for (int i = 0; i < m; ++i) { const int c = (int)info.col_ind[i]; fwrite(&c, sizeof(int), 1, o); }
    fwrite(info.start_ring_index.data(), sizeof(int), R, o);
    fwrite(info.end_ring_index.data(), sizeof(int), R, o);
    const float ori[3] = {info.start_orientation, info.end_orientation, info.orientation_diff};
    fwrite(ori, sizeof(float), 3, o);
    fclose(o);
    return 0;
}
