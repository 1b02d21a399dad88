// Drives include/superodom_b200/FeatureExtraction.hpp the way the feature-extraction node does:
//   removePointDistortion(lidar_start_time, lidar_end_time, buffer, lidar_msg); uniformFeatureExtraction(lidar_msg, plannerPoints, skip, min_range)
// Input: binary file from tests/test_cpp_shim.py; output: the deskewed x,y,z of every point, then the extracted cloud, as raw floats.
#include <cstdio>
#include <map>
#include <memory>
#include <vector>

#include "superodom_b200/FeatureExtraction.hpp"

namespace so = super_odometry::b200;

struct PointXYZITR {        // point_os::PointcloudXYZITR layout: x,y,z,pad | intensity, time, ring | pad -> 32 bytes, time at byte 20
    float x, y, z, pad0;
    float intensity, time;
    uint16_t ring, pad1;
    float pad2;
};
struct PointXYZI { float x, y, z, pad0; float intensity, pad1[3]; };
template <class P> struct CloudT {
    std::vector<P> points;
    size_t size() const { return points.size(); }
};
struct Odom { double q[4]; double p[3]; };      // stands in for nav_msgs::msg::Odometry::SharedPtr

static bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

int main(int argc, char** argv) {
    static_assert(sizeof(PointXYZITR) == 32, "layout");
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t n = 0, m = 0, imu_only = 0, skip = 0;
    double start_time = 0, til[7];
    float block_range = 0;
    if (!rd(f, &n, 4) || !rd(f, &m, 4) || !rd(f, &imu_only, 4) || !rd(f, &skip, 4) || !rd(f, &block_range, 4) || !rd(f, &start_time, 8) || !rd(f, til, 56)) return 2;
    auto raw = std::make_shared<CloudT<PointXYZITR>>();
    raw->points.resize(n);
    for (int i = 0; i < n; ++i) {
        float v[8];
        if (!rd(f, v, 32)) return 2;
        PointXYZITR& p = raw->points[i];
        p.x = v[0]; p.y = v[1]; p.z = v[2]; p.pad0 = 1.f; p.intensity = v[4]; p.time = v[5]; p.ring = uint16_t(v[6]); p.pad1 = 0; p.pad2 = 0.f;
    }
    std::map<double, std::shared_ptr<Odom>> buffer;      // MapRingBuffer::measMap_
    for (int k = 0; k < m; ++k) {
        double t, pose[7];
        if (!rd(f, &t, 8) || !rd(f, pose, 56)) return 2;
        auto o = std::make_shared<Odom>();
        o->p[0] = pose[0]; o->p[1] = pose[1]; o->p[2] = pose[2];
        o->q[0] = pose[3]; o->q[1] = pose[4]; o->q[2] = pose[5]; o->q[3] = pose[6];
        buffer[t] = o;
    }
    fclose(f);
    so::Context ctx;
    ctx.cfg.max_scan_points = 1u << 18;
    ctx.cfg.max_map_points = 1024;
    try {
        ctx.ensure(0.1f, 0.2f);
        auto extract = [](const std::shared_ptr<Odom>& d) {
            return so::Transformd(so::Quaterniond(d->q[3], d->q[0], d->q[1], d->q[2]), so::Vector3d(d->p[0], d->p[1], d->p[2]));
        };
        const so::Transformd start = so::removePointDistortion(ctx, start_time, start_time + 0.1, buffer, extract, imu_only != 0, so::Transformd::from_pose7(til), raw);
        auto planner = std::make_shared<CloudT<PointXYZI>>();
        so::uniformFeatureExtraction(ctx, raw, planner, skip, block_range);
        FILE* o = fopen(argv[2], "wb");
        if (!o) return 2;
        double s7[7];
        start.to_pose7(s7);
        fwrite(s7, 8, 7, o);
        for (const auto& p : raw->points) { const float v[3] = {p.x, p.y, p.z}; fwrite(v, 4, 3, o); }
        const int32_t k = int32_t(planner->size());
        fwrite(&k, 4, 1, o);
        for (const auto& p : planner->points) { const float v[4] = {p.x, p.y, p.z, p.intensity}; fwrite(v, 4, 4, o); }
        fclose(o);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
