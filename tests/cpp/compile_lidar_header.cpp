// compile-only check of include/lvio_b200/lidar_features.h against a PCL-shaped point type (tests/test_capi_cpu.py)
#include <vector>
#include "lvio_b200/lidar_features.h"
struct alignas(16) PointXYZI { float x, y, z, pad0, intensity, pad1, pad2, pad3; };   // sizeof == 32 like pcl::PointXYZI
struct PointXYZ { float x, y, z, pad; };
int main() {
    lvb::LidarFeatureExtractor ex(64, 1800, 0.427, 24.9, 60, 0.1036, 5, 30, 0.2);
    std::vector<PointXYZ> scan;
    std::vector<PointXYZI> ground, surf;
    const bool ok = ex.Process(scan, ground, surf);
    std::vector<lvb::ImuInterval> iv;
    std::vector<double> out;
    const double noise[4] = {0.1, 0.01, 0.001, 0.0001};
    return (ok ? 0 : 1) + (lvb::preintegrate_batch(iv, noise, out) ? 0 : 2);
}
