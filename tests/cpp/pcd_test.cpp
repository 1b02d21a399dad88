// Reads a PCD file with include/superodom_b200/PcdIO.hpp the way laserMapping does in localization mode
// (utils::readPointCloud(config_.map_dir, laserCloudPrior), laserMapping.cpp:163-166) and dumps the points as raw floats.
#include <cstdio>
#include <memory>
#include <vector>

#include "superodom_b200/PcdIO.hpp"

struct PointXYZI { float x, y, z, pad0; float intensity, pad1[3]; };
struct Cloud { std::vector<PointXYZI> points; size_t size() const { return points.size(); } };

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    auto cloud = std::make_shared<Cloud>();
    if (!super_odometry::b200::readPointCloud(argv[1], cloud)) return 1;
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 2;
    for (const auto& p : cloud->points) { const float v[4] = {p.x, p.y, p.z, p.intensity}; fwrite(v, 4, 4, o); }
    fclose(o);
    return 0;
}
