// =============================================================================
// superodom_b200/LidarSlam.hpp -- header-only C++ shim that keeps the reference's
// LidarSLAM / LocalMap call surface (super_odometry/include/super_odometry/
// LidarProcess/LidarSlam.h, LocalMap.h) and forwards to the C ABI
// (superodom_b200.h).  laserMapping.cpp uses exactly these members
// (SURVEY.md section 8b):
//   call   slam.Localization(bool, PredictionSource, Transformd, edge, planner, t)   laserMapping.cpp:713-714
//   write  slam.localMap.{lineRes_,planeRes_}, Visual_confidence_factor, Pos/Ori_degeneracy_threshold,
//          LocalizationICPMaxIter, OptSet.*, map_dir, localization_mode, init_*, last_T_w_lidar,
//          frame_count, laser_imu_sync                                               :102-120,312,648-649,703-711,740-741
//          slam.localMap.setOrigin(t), slam.localMap.addSurfPointCloud(cloud)        :161,166
//   read   slam.T_w_lidar, startupCount, isDegenerate, stats, pos_in_localmap,
//          localMap.get5x5LocalMap(pos), localMap.getAllLocalMap()                   :387,439,450,563,581-596,734-738
//
// The shim is dependency-free: it ships tiny Eigen-compatible value types
// (x()/y()/z()/w() accessors, converting constructors from anything with the same
// accessors, so Eigen::Quaterniond / Eigen::Vector3d / the reference's Transformd
// drop in) and is templated on the point-cloud pointer type (anything with
// ->points.data(), ->size() and a point struct holding float x,y,z,intensity --
// pcl::PointCloud<pcl::PointXYZI>::Ptr in the node).  See INTEGRATION.md.
// =============================================================================
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../superodom_b200.h"

namespace super_odometry {
namespace b200 {

struct Vector3d {
    double v[3] = {0, 0, 0};
    Vector3d() = default;
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    template <class E, class = decltype(std::declval<const E&>().x())>
    Vector3d(const E& e) : v{double(e.x()), double(e.y()), double(e.z())} {}
    double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; }
    double& x() { return v[0]; } double& y() { return v[1]; } double& z() { return v[2]; }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
struct Vector3i {
    int v[3] = {0, 0, 0};
    int x() const { return v[0]; } int y() const { return v[1]; } int z() const { return v[2]; }
    int& x() { return v[0]; } int& y() { return v[1]; } int& z() { return v[2]; }
};
struct Quaterniond {
    double q[4] = {0, 0, 0, 1};   // x y z w (Eigen coefficient order)
    Quaterniond() = default;
    Quaterniond(double w, double x, double y, double z) : q{x, y, z, w} {}   // Eigen ctor order (w, x, y, z)
    template <class E, class = decltype(std::declval<const E&>().w())>
    Quaterniond(const E& e) : q{double(e.x()), double(e.y()), double(e.z()), double(e.w())} {}
    double x() const { return q[0]; } double y() const { return q[1]; } double z() const { return q[2]; } double w() const { return q[3]; }
    Quaterniond conjugate() const { return Quaterniond(q[3], -q[0], -q[1], -q[2]); }
    Quaterniond operator*(const Quaterniond& b) const {
        return Quaterniond(q[3] * b.q[3] - q[0] * b.q[0] - q[1] * b.q[1] - q[2] * b.q[2],
                           q[3] * b.q[0] + q[0] * b.q[3] + q[1] * b.q[2] - q[2] * b.q[1],
                           q[3] * b.q[1] + q[1] * b.q[3] + q[2] * b.q[0] - q[0] * b.q[2],
                           q[3] * b.q[2] + q[2] * b.q[3] + q[0] * b.q[1] - q[1] * b.q[0]);
    }
    Vector3d operator*(const Vector3d& p) const {
        double ux = q[1] * p.v[2] - q[2] * p.v[1], uy = q[2] * p.v[0] - q[0] * p.v[2], uz = q[0] * p.v[1] - q[1] * p.v[0];
        ux += ux; uy += uy; uz += uz;
        return Vector3d(p.v[0] + q[3] * ux + (q[1] * uz - q[2] * uy), p.v[1] + q[3] * uy + (q[2] * ux - q[0] * uz),
                        p.v[2] + q[3] * uz + (q[0] * uy - q[1] * ux));
    }
};
// utils/Twist.h `Transformd`: rot + pos, operator*, inverse()
struct Transformd {
    Quaterniond rot;
    Vector3d pos;
    Transformd() = default;
    Transformd(const Quaterniond& r, const Vector3d& p) : rot(r), pos(p) {}
    template <class T, class = decltype(std::declval<const T&>().rot.w())>
    Transformd(const T& t) : rot(t.rot), pos(t.pos) {}
    Transformd inverse() const {
        Transformd o; o.rot = rot.conjugate(); Vector3d t = o.rot * pos; o.pos = Vector3d(-t.x(), -t.y(), -t.z()); return o;
    }
    Transformd operator*(const Transformd& b) const {
        Transformd o; o.rot = rot * b.rot; Vector3d t = rot * b.pos; o.pos = Vector3d(t.x() + pos.x(), t.y() + pos.y(), t.z() + pos.z()); return o;
    }
    void to_pose7(double out[7]) const { out[0] = pos.x(); out[1] = pos.y(); out[2] = pos.z(); out[3] = rot.x(); out[4] = rot.y(); out[5] = rot.z(); out[6] = rot.w(); }
    static Transformd from_pose7(const double p[7]) { return Transformd(Quaterniond(p[6], p[3], p[4], p[5]), Vector3d(p[0], p[1], p[2])); }
};

// super_odometry_msgs/msg/{IterationStats,OptimizationStats}.msg as plain structs (same field names)
struct IterationStats { double translation_norm = 0, rotation_norm = 0, num_surf_from_scan = 0, num_corner_from_scan = 0; };
struct OptimizationStats {
    int32_t laser_cloud_surf_from_map_num = 0, laser_cloud_corner_from_map_num = 0, laser_cloud_surf_stack_num = 0, laser_cloud_corner_stack_num = 0;
    double total_translation = 0, total_rotation = 0, translation_from_last = 0, rotation_from_last = 0, time_elapsed = 0, latency = 0;
    int32_t n_iterations = 0;
    double average_distance = 0, uncertainty_x = 0, uncertainty_y = 0, uncertainty_z = 0, uncertainty_roll = 0, uncertainty_pitch = 0, uncertainty_yaw = 0;
    int32_t plane_match_success = 0, plane_no_enough_neighbor = 0, plane_neighbor_too_far = 0, plane_badpca_structure = 0,
            plane_invalid_numerical = 0, plane_mse_too_large = 0, plane_unknown = 0, prediction_source = 0;
    std::vector<IterationStats> iterations;
};

class Error : public std::runtime_error { public: using std::runtime_error::runtime_error; };

// Shared handle to the device context; LocalMap and LidarSLAM both forward through it.
struct Context {
    so_ctx* h = nullptr;
    so_config cfg{};
    void ensure(float line_res, float plane_res) {
        if (!h) {
            cfg.line_res = line_res; cfg.plane_res = plane_res;
            h = so_create(&cfg);
            if (!h) throw Error(std::string("so_create: ") + so_last_error());
        }
    }
    ~Context() { if (h) so_destroy(h); }
};

// point layout of the caller's cloud type
template <class PointT>
inline void point_layout(size_t* stride, size_t* ioff) {
    PointT p{};
    *stride = sizeof(PointT);
    *ioff = size_t(reinterpret_cast<const char*>(&p.intensity) - reinterpret_cast<const char*>(&p));
}

// ------------------------------------------------------------------------------------------------ LocalMap (LocalMap.h:123-766)
class LocalMap {
public:
    static constexpr int laserCloudWidth = 21, laserCloudHeight = 21, laserCloudDepth = 11;
    static constexpr int laserCloudNum = laserCloudWidth * laserCloudHeight * laserCloudDepth;
    float lineRes_ = 0.2f;     // LocalMap.h:761
    float planeRes_ = 0.4f;    // LocalMap.h:762
    Vector3i origin_;

    explicit LocalMap(Context* c) : ctx_(c) { origin_.v[0] = 10; origin_.v[1] = 10; origin_.v[2] = 5; }

    template <class V> Vector3i setOrigin(const V& t_w_cur) {                       // LocalMap.h:146-164
        sync_resolution();
        const double t[3] = {double(t_w_cur.x()), double(t_w_cur.y()), double(t_w_cur.z())};
        check(so_map_set_origin(ctx_->h, t, origin_.v), "so_map_set_origin");
        return origin_;
    }
    template <class V> Vector3i shiftMap(const V& t_w_cur) {                        // LocalMap.h:169-287
        sync_resolution();
        const double t[3] = {double(t_w_cur.x()), double(t_w_cur.y()), double(t_w_cur.z())};
        Vector3i ijk;
        check(so_map_shift(ctx_->h, t, ijk.v), "so_map_shift");
        check(so_map_get_origin(ctx_->h, origin_.v), "so_map_get_origin");
        return ijk;
    }
    std::tuple<int, int> get5x5LocalMapFeatureSize(const Vector3i& position) {      // LocalMap.h:291-318
        sync_resolution();
        int32_t ne = 0, ns = 0;
        check(so_map_counts_5x5(ctx_->h, position.v, &ne, &ns), "so_map_counts_5x5");
        return std::make_tuple(int(ne), int(ns));
    }
    template <class Cloud> void addSurfPointCloud(Cloud& laserCloudSurfStack) {     // LocalMap.h:591-645 (world-frame points)
        sync_resolution();
        if (laserCloudSurfStack.size() == 0) return;
        size_t stride, ioff;
        point_layout<typename std::decay<decltype(laserCloudSurfStack.points[0])>::type>(&stride, &ioff);
        check(so_map_add_surf(ctx_->h, laserCloudSurfStack.points.data(), laserCloudSurfStack.size(), stride, ioff), "so_map_add_surf");
    }
    template <class Cloud> void addEdgePointCloud(Cloud& laserCloudEdgeStack) {     // LocalMap.h:529-589 (world-frame points)
        sync_resolution();
        if (laserCloudEdgeStack.size() == 0) return;
        size_t stride, ioff;
        point_layout<typename std::decay<decltype(laserCloudEdgeStack.points[0])>::type>(&stride, &ioff);
        check(so_map_add_edge(ctx_->h, laserCloudEdgeStack.points.data(), laserCloudEdgeStack.size(), stride, ioff), "so_map_add_edge");
    }
    // getAllLocalMap / get5x5LocalMap (LocalMap.h:647-687) fill a caller cloud type (resize + x,y,z,intensity)
    template <class Cloud> Cloud getAllLocalMap() { return download<Cloud>(0, nullptr); }
    template <class Cloud> Cloud get5x5LocalMap(const Vector3i& position) { return download<Cloud>(1, position.v); }
    size_t size() const { return ctx_->h ? so_map_size(ctx_->h) : 0; }

    void sync_resolution() {
        ctx_->ensure(lineRes_, planeRes_);
        if (lineRes_ != pushed_line_ || planeRes_ != pushed_plane_) {
            check(so_map_set_resolution(ctx_->h, lineRes_, planeRes_), "so_map_set_resolution");
            pushed_line_ = lineRes_; pushed_plane_ = planeRes_;
        }
    }

private:
    template <class Cloud> Cloud download(int mode, const int32_t* ijk) {
        sync_resolution();
        size_t n = 0;
        const size_t cap = so_map_size(ctx_->h);
        std::vector<float> buf(4 * (cap ? cap : 1));
        check(so_map_download(ctx_->h, mode, ijk, buf.data(), cap, &n), "so_map_download");
        Cloud out;
        out.points.resize(n);
        for (size_t i = 0; i < n; ++i) { auto& p = out.points[i]; p.x = buf[4 * i]; p.y = buf[4 * i + 1]; p.z = buf[4 * i + 2]; p.intensity = buf[4 * i + 3]; }
        return out;
    }
    static void check(int rc, const char* what) { if (rc < 0) throw Error(std::string(what) + ": " + so_last_error()); }
    Context* ctx_;
    float pushed_line_ = -1.f, pushed_plane_ = -1.f;
};

// ------------------------------------------------------------------------------------------------ LidarSLAM (LidarSlam.h:40-423)
class LidarSLAM {
public:
    enum class PredictionSource { IMU_ORIENTATION, LIO_ODOM, VIO_ODOM, NEURAL_IMU_ODOM, CONSTANT_VELOCITY };
    enum MatchingResult : uint8_t { SUCCESS = 0, NOT_ENOUGH_NEIGHBORS = 1, NEIGHBORS_TOO_FAR = 2, BAD_PCA_STRUCTURE = 3,
                                    INVAVLID_NUMERICAL = 4, MSE_TOO_LARGE = 5, UNKNON = 6, nRejectionCauses = 7 };
    enum Feature_observability : uint8_t { rx_cross = 0, neg_rx_cross = 1, ry_cross = 2, neg_ry_cross = 3, rz_cross = 4, neg_rz_cross = 5,
                                           tx_dot = 6, ty_dot = 7, tz_dot = 8, nFeatureObs = 9 };
    struct LaserOptSet {
        double imu_roll_pitch[4] = {0, 0, 0, 1};
        bool debug_view_enabled = false, use_imu_roll_pitch = false;
        float velocity_failure_threshold = 30.f, yaw_ratio = 0.f;
        int max_surface_features = 2000;
    };
    struct RegistrationError {                                   // LidarSlam.h:122-148
        double PositionError = 0., PositionUncertainty = 0., MaxPositionError = 0.1, PosInverseConditionNum = 1.0;
        Vector3d PositionErrorDirection;
        double OrientationError = 0., OrientationUncertainty = 0., MaxOrientationError = 10, OriInverseConditionNum = 1.0;
        Vector3d OrientationErrorDirection;
        std::array<double, 36> Covariance{};                     // row-major, DoF order X,Y,Z,rX,rY,rZ
    };
    struct LidarOdomUncertainty { double uncertainty_x = 0, uncertainty_y = 0, uncertainty_z = 0, uncertainty_roll = 0, uncertainty_pitch = 0, uncertainty_yaw = 0; };

    Context context;                 // owns the device context (declared first: localMap points into it)
    LocalMap localMap;
    OptimizationStats stats;
    RegistrationError LocalizationUncertainty;
    LidarOdomUncertainty lidarOdomUncer;
    LaserOptSet OptSet;
    Transformd T_w_lidar, last_T_w_lidar, T_w_initial_guess;
    Vector3i pos_in_localmap;
    int frame_count = 0, laser_imu_sync = 0, startupCount = 0;
    float Pos_degeneracy_threshold = 0, Ori_degeneracy_threshold = 0, Visual_confidence_factor = 0;
    std::string map_dir;
    float init_x = 0, init_y = 0, init_z = 0, init_roll = 0, init_pitch = 0, init_yaw = 0, localization_mode = 0, update_map = 0;
    double lasttimeLaserOdometry = 0;
    bool bInitialization = false, isDegenerate = false;          // isDegenerate is never set by the reference (LidarSlam.cpp:976-985)
    std::array<int, nFeatureObs> PlaneFeatureHistogramObs{};
    std::array<int, nRejectionCauses> MatchRejectionHistogramLine{}, MatchRejectionHistogramPlane{};
    size_t LocalizationICPMaxIter = 4;                           // LidarSlam.h:273
    so_icp_result last_result{};                                 // everything the device reported for the last scan

    explicit LidarSLAM(int device = 0, uint32_t max_map_points = 0, uint32_t max_scan_points = 0) : localMap(&context) {
        context.cfg.device = device; context.cfg.max_map_points = max_map_points; context.cfg.max_scan_points = max_scan_points;
    }
    template <class NodePtr> void initROSInterface(NodePtr) {}   // the six Float32 uncertainty publishers stay in the node (LidarSlam.cpp:20-28)

    // LidarSLAM::Localization (LidarSlam.cpp:30-51)
    template <class CloudPtr>
    void Localization(bool initialization, PredictionSource predictodom, const Transformd& position, const CloudPtr& edge_point,
                      const CloudPtr& planner_point, double timeLaserOdometry) {
        T_w_lidar = position; T_w_initial_guess = position; last_T_w_lidar = position;          // initializeState (:53-57)
        localMap.sync_resolution();
        size_t stride, ioff;
        using PointT = typename std::decay<decltype(planner_point->points[0])>::type;
        point_layout<PointT>(&stride, &ioff);
        const void* surf = planner_point->size() ? planner_point->points.data() : nullptr;
        double pose[7];
        if (!initialization) {                                                                   // initializeMapping (:83-94)
            localMap.setOrigin(T_w_lidar.pos);
            T_w_lidar.to_pose7(pose);
            if (edge_point && edge_point->size()) check(so_map_add_scan_edge(context.h, edge_point->points.data(), edge_point->size(), stride, ioff, pose), "so_map_add_scan_edge");
            if (planner_point->size()) check(so_map_add_scan(context.h, surf, planner_point->size(), stride, ioff, pose), "so_map_add_scan");
            lasttimeLaserOdometry = timeLaserOdometry;
            return;
        }
        EstimateLidarUncertainty();                                                              // (:915-986), from the PREVIOUS scan's histogram
        so_icp_opts o{};
        o.max_icp_iters = int32_t(LocalizationICPMaxIter);
        o.max_surface_features = OptSet.max_surface_features;
        o.lm_max_iterations = 4;
        o.yaw_ratio = OptSet.yaw_ratio;
        // shouldAddAbsolutePoseConstraints (LidarSlam.cpp:281-283)
        if (predictodom == PredictionSource::VIO_ODOM && isDegenerate && Visual_confidence_factor != 0) {
            o.use_pose_prior = 1;
            o.visual_confidence_factor = Visual_confidence_factor;
            o.prior_uncertainty[0] = float(lidarOdomUncer.uncertainty_x);
            o.prior_uncertainty[1] = float(lidarOdomUncer.uncertainty_y);
            o.prior_uncertainty[2] = float(lidarOdomUncer.uncertainty_z);
        }
        position.to_pose7(pose);
        const int rc = so_register(context.h, surf, planner_point->size(), edge_point ? (const void*)edge_point->points.data() : nullptr,
                                   edge_point ? edge_point->size() : 0, stride, ioff, pose, &o, &last_result);
        check(rc, "so_register");
        const so_icp_result& r = last_result;
        pos_in_localmap.v[0] = r.pos_in_localmap[0]; pos_in_localmap.v[1] = r.pos_in_localmap[1]; pos_in_localmap.v[2] = r.pos_in_localmap[2];
        check(so_map_get_origin(context.h, localMap.origin_.v), "so_map_get_origin");
        stats.laser_cloud_corner_from_map_num = r.map_edge_5x5; stats.laser_cloud_surf_from_map_num = r.map_surf_5x5;   // updateFeatureStats (:371-377)
        stats.laser_cloud_corner_stack_num = int32_t(edge_point ? edge_point->size() : 0); stats.laser_cloud_surf_stack_num = int32_t(planner_point->size());
        stats.iterations.clear();
        if (rc == SO_STATUS_NOT_ENOUGH_FEATURES) return;                                         // (:113-116) WARN + return, pose = prior
        for (int i = 0; i < r.n_iterations; ++i) {                                               // recordIterationStats (:242-251)
            IterationStats it; it.num_surf_from_scan = r.iter_n_surf[i]; it.num_corner_from_scan = r.iter_n_edge[i];
            it.translation_norm = r.iter_dtrans[i]; it.rotation_norm = r.iter_drot[i];
            stats.iterations.push_back(it);
        }
        stats.prediction_source = r.prediction_source;                                            // (:278,297)
        for (int i = 0; i < 9; ++i) PlaneFeatureHistogramObs[i] = r.hist_obs[i];
        for (int i = 0; i < 7; ++i) { MatchRejectionHistogramPlane[i] = r.hist_reject_plane[i]; MatchRejectionHistogramLine[i] = r.hist_reject_line[i]; }
        LocalizationUncertainty.PositionError = r.pos_err; LocalizationUncertainty.PosInverseConditionNum = r.pos_inv_cond;
        LocalizationUncertainty.OrientationError = r.ori_err_deg; LocalizationUncertainty.OriInverseConditionNum = r.ori_inv_cond;
        LocalizationUncertainty.PositionErrorDirection = Vector3d(r.pos_dir[0], r.pos_dir[1], r.pos_dir[2]);
        LocalizationUncertainty.OrientationErrorDirection = Vector3d(r.ori_dir[0], r.ori_dir[1], r.ori_dir[2]);
        std::memcpy(LocalizationUncertainty.Covariance.data(), r.cov, sizeof(r.cov));
        T_w_lidar = Transformd::from_pose7(r.pose);                                              // after MannualYawCorrection (:891-913)
        stats.time_elapsed = r.time_ms;                                                          // updateOptimizationStats (:198-210)
        stats.total_translation = r.total_translation; stats.total_rotation = r.total_rotation;
        stats.translation_from_last = r.translation_from_last; stats.rotation_from_last = r.rotation_from_last;
        stats.n_iterations = r.n_iterations;
        last_T_w_lidar = T_w_lidar;
        // checkMotionThresholds always accepts (:193) -> transformAndAddToMap (:163-167)
        if (edge_point && edge_point->size()) check(so_map_add_scan_edge(context.h, edge_point->points.data(), edge_point->size(), stride, ioff, r.pose), "so_map_add_scan_edge");
        if (planner_point->size()) check(so_map_add_registered_scan(context.h, r.pose), "so_map_add_registered_scan");   // the scan is still on the device
        lasttimeLaserOdometry = timeLaserOdometry;
    }

    // LidarSLAM::EstimateLidarUncertainty (LidarSlam.cpp:915-975)
    void EstimateLidarUncertainty() {
        const auto& h = PlaneFeatureHistogramObs;
        const double tt = double(h[6]) + double(h[7]) + double(h[8]);
        const double tr = double(h[0]) + h[1] + h[2] + h[3] + h[4] + h[5];
        auto cap = [](double v) { return v < 1.0 ? v : 1.0; };
        lidarOdomUncer.uncertainty_x = cap(h[6] / tt * 3); lidarOdomUncer.uncertainty_y = cap(h[7] / tt * 3); lidarOdomUncer.uncertainty_z = cap(h[8] / tt * 3);
        lidarOdomUncer.uncertainty_roll = cap((double(h[0]) + h[1]) / tr * 3); lidarOdomUncer.uncertainty_pitch = cap((double(h[2]) + h[3]) / tr * 3);
        lidarOdomUncer.uncertainty_yaw = cap((double(h[4]) + h[5]) / tr * 3);
        if (tt == 0 || tr == 0) lidarOdomUncer = LidarOdomUncertainty{};
        stats.uncertainty_x = lidarOdomUncer.uncertainty_x; stats.uncertainty_y = lidarOdomUncer.uncertainty_y; stats.uncertainty_z = lidarOdomUncer.uncertainty_z;
        stats.uncertainty_roll = lidarOdomUncer.uncertainty_roll; stats.uncertainty_pitch = lidarOdomUncer.uncertainty_pitch; stats.uncertainty_yaw = lidarOdomUncer.uncertainty_yaw;
    }

private:
    static void check(int rc, const char* what) { if (rc < 0) throw Error(std::string(what) + ": " + so_last_error()); }
};

}  // namespace b200
}  // namespace super_odometry
