/* =============================================================================
 * superodom_b200.h -- C ABI of the B200-native ICP registration hot path.
 *
 * This is the drop-in boundary for ONE path of superxslam/SuperOdom: the per-scan
 * registration in LidarSLAM::performLocalizationAndMapping.  The reference has no
 * FFI layer; its seam is the C++ class surface of LidarSLAM / LocalMap as used by
 * laserMapping.  Each entry point below names the reference
 * interface it replaces (file:line under /root/reference/super_odometry/).  The C++
 * shim include/superodom_b200/LidarSlam.hpp keeps the reference member names and
 * forwards to these functions; INTEGRATION.md shows the maintainer-side binding.
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every host buffer; the context owns
 *     all device memory, streams and CUDA graphs; no allocation crosses the ABI.
 *   - pose7 = {tx,ty,tz,qx,qy,qz,qw}: the reference's pose_parameters layout
 *     (src/LidarProcess/LidarSlam.cpp:7-9).
 *   - point clouds are arrays of points with a byte stride; x,y,z are floats at byte
 *     offsets 0,4,8 and intensity is a float at byte offset `intensity_offset`
 *     (pcl::PointXYZI: stride 32, intensity at 16; packed float4: stride 16, at 12).
 *   - return value: 0 = ok; >0 = reference-defined soft status (see SO_STATUS_*);
 *     <0 = CUDA / argument error, text via so_last_error().  Nothing throws.
 *   - one host thread per context, one context per GPU (the reference is
 *     non-reentrant: file-scope pose globals, LidarSlam.cpp:7-9).
 *   - there is NO CPU fallback: every entry point fails with SO_ERR_CUDA when no
 *     sm_100 device / kernel image is available.
 * ============================================================================= */
#ifndef SUPERODOM_B200_H
#define SUPERODOM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_MAX_ICP_ITERS 32

/* soft statuses (>0) and errors (<0) */
#define SO_OK 0
#define SO_STATUS_NOT_ENOUGH_FEATURES 1 /* hasEnoughFeatures() false: pose = prior (LidarSlam.cpp:113-116,379-381) */
#define SO_STATUS_NO_CORRESPONDENCES 2  /* every scan point was rejected; reference would CHECK-fail in ceres::Covariance */
#define SO_ERR_ARG (-1)
#define SO_ERR_CUDA (-2)
#define SO_ERR_CAPACITY (-3)

/* LidarSLAM::MatchingResult (include/super_odometry/LidarProcess/LidarSlam.h:85-94) */
enum so_match_result {
    SO_MATCH_SUCCESS = 0, SO_MATCH_NOT_ENOUGH_NEIGHBORS = 1, SO_MATCH_NEIGHBORS_TOO_FAR = 2,
    SO_MATCH_BAD_PCA_STRUCTURE = 3, SO_MATCH_INVALID_NUMERICAL = 4, SO_MATCH_MSE_TOO_LARGE = 5,
    SO_MATCH_UNKNOWN = 6, SO_MATCH_N = 7,
    SO_MATCH_SKIPPED = 255 /* not in the reference: point dropped by shouldProcessPoint (LidarSlam.cpp:353-359) */
};

typedef struct so_ctx so_ctx;

/* Construction-time capacities.  Zero fields take defaults. */
typedef struct {
    int32_t device;            /* CUDA device ordinal */
    uint32_t max_map_points;   /* default 4,194,304 */
    uint32_t max_scan_points;  /* per scan, default 262,144 */
    uint32_t max_batch;        /* scans registered concurrently by so_register_batch*, default 1 */
    float line_res;            /* LocalMap::lineRes_  (LocalMap.h:761; node default laserMapping.cpp:102-103) */
    float plane_res;           /* LocalMap::planeRes_ (LocalMap.h:762) */
} so_config;

/* Per-call options == the LidarSLAM members laserMapping writes (laserMapping.cpp:102-120,703-711). */
typedef struct {
    int32_t max_icp_iters;        /* LidarSLAM::LocalizationICPMaxIter (LidarSlam.h:273); config max_iterations */
    int32_t max_surface_features; /* OptSet.max_surface_features (LidarSlam.h:118); 0 = uncapped */
    int32_t lm_max_iterations;    /* ceres options.max_num_iterations = 4 (LidarSlam.cpp:232); 0 -> 4 */
    float yaw_ratio;              /* OptSet.yaw_ratio (LidarSlam.cpp:905) */
    int32_t skip_map_checks;      /* 1: do not run shiftMap / hasEnoughFeatures (replay against a frozen map) */
    /* SE3AbsolutatePoseFactor on T_w_initial_guess (LidarSlam.cpp:281-298): set when shouldAddAbsolutePoseConstraints()
     * holds, i.e. predictodom == VIO_ODOM && isDegenerate && Visual_confidence_factor != 0 (dormant in the shipped node:
     * isDegenerate is never set, LidarSlam.cpp:976-985). */
    int32_t use_pose_prior;
    float visual_confidence_factor;   /* LidarSLAM::Visual_confidence_factor */
    float prior_uncertainty[3];       /* lidarOdomUncer.uncertainty_{x,y,z} (EstimateLidarUncertainty, LidarSlam.cpp:915-964) */
} so_icp_opts;

/* Everything the reference leaves in LidarSLAM members after Localization():
 * T_w_lidar, stats (OptimizationStats.msg), LocalizationUncertainty (RegistrationError, LidarSlam.h:122-148),
 * PlaneFeatureHistogramObs / MatchRejectionHistogram* (LidarSlam.h:267-269). */
typedef struct {
    double pose[7];       /* T_w_lidar after MannualYawCorrection (LidarSlam.cpp:891-913) */
    double pose_opt[7];   /* optimiser output (pose_parameters) before the RPY round trip */
    int32_t status;       /* SO_OK / SO_STATUS_* */
    int32_t n_iterations; /* ICP iterations executed (stats.iterations.size()) */
    int32_t iter_n_surf[SO_MAX_ICP_ITERS];        /* IterationStats.num_surf_from_scan */
    int32_t iter_n_edge[SO_MAX_ICP_ITERS];        /* IterationStats.num_corner_from_scan */
    double iter_dtrans[SO_MAX_ICP_ITERS];         /* IterationStats.translation_norm */
    double iter_drot[SO_MAX_ICP_ITERS];           /* IterationStats.rotation_norm */
    int32_t iter_lm_steps[SO_MAX_ICP_ITERS];      /* ceres summary: iterations of this solve (diagnostic) */
    int32_t iter_lm_successful[SO_MAX_ICP_ITERS]; /* ceres summary.num_successful_steps (drives the break, :141) */
    int32_t iter_lm_termination[SO_MAX_ICP_ITERS];/* 0 max-iter 1 gradient 2 parameter 3 function 4 radius 5 invalid 6 empty */
    double iter_cost[SO_MAX_ICP_ITERS];           /* final cost of this solve */
    int32_t hist_obs[9];          /* PlaneFeatureHistogramObs of the last ICP iteration */
    int32_t hist_reject_plane[7]; /* MatchRejectionHistogramPlane of the last ICP iteration */
    int32_t hist_reject_line[7];  /* MatchRejectionHistogramLine of the last ICP iteration (all 0 without an edge cloud) */
    double cov[36];               /* RegistrationError::Covariance, row-major, order x,y,z,rx,ry,rz */
    double pos_err, pos_dir[3], pos_inv_cond;     /* PositionError, PositionErrorDirection, PosInverseConditionNum */
    double ori_err_deg, ori_dir[3], ori_inv_cond; /* OrientationError (degrees), ...Direction, OriInverseConditionNum */
    double total_translation, total_rotation, translation_from_last, rotation_from_last; /* stats.* (LidarSlam.cpp:198-210) */
    int32_t map_surf_5x5, map_edge_5x5, scan_surf_num, scan_edge_num; /* stats.laser_cloud_* (LidarSlam.cpp:371-377) */
    int32_t pos_in_localmap[3];   /* LidarSLAM::pos_in_localmap (shiftMap return) */
    int32_t prediction_source;    /* stats.prediction_source: 1 when the pose prior rows were added (LidarSlam.cpp:278,297) */
    double time_ms;               /* stats.time_elapsed: the ICP loop, device time measured with CUDA events */
    double time_total_ms;         /* whole so_register call incl. H2D / D2H (host clock) */
    int32_t knn_searched;         /* telemetry, summed over the ICP iterations: queries that ran the neighbour search ... */
    int32_t knn_verified;         /* ... and queries whose previous five neighbours were proven to still be their 5-NN (no search) */
} so_icp_result;

/* One accepted/rejected plane correspondence, for stage-level parity tests
 * (LidarSLAM::OptimizationParameter, LidarSlam.h:209-222, reduced to what the solver reads). */
typedef struct {
    double n[3];      /* NormDir */
    double d;         /* negative_OA_dot_norm */
    double w;         /* residualCoefficient */
    uint32_t nn[5];   /* neighbour ids = index of the map point in the order given to so_map_set_points / so_map_add_surf output */
    float nn_d2[5];
    uint8_t status;   /* so_match_result */
    uint8_t obs[3];   /* Feature_observability labels histogrammed (LidarSlam.cpp:336-339) */
    uint8_t pad_[4];
} so_corr;

/* One edge / line correspondence (stage-level parity of the dormant edge branch, LidarSlam.cpp:402-493). */
typedef struct {
    double a[3], b[3];      /* corres: mean +- 0.1 * line direction */
    double w;               /* residualCoefficient */
    uint32_t nn[10];        /* the 10 nearest edge-map points (ids as given to so_map_set_edge_points / insert output order) */
    uint32_t selected_mask; /* bit j: nn[j] kept by nearestKSearchSpecificEdgePoint's best-line selection (bit 0 always) */
    uint8_t status;         /* so_match_result */
    uint8_t n_selected;
    uint8_t pad_[2];
} so_edge_corr;

/* ---- lifetime ------------------------------------------------------------------------------------ */
/* replaces: LidarSLAM::LidarSLAM() + LocalMap::LocalMap() (LidarSlam.cpp:14-19, LocalMap.h:141-144) */
so_ctx* so_create(const so_config* cfg);
void so_destroy(so_ctx* ctx);
const char* so_last_error(void);
/* 1 when a CUDA device of compute capability 10.x is usable by this library */
int so_device_available(void);
/* Use an externally owned CUDA stream (cudaStream_t as void*) for all work of this context; NULL = own stream. */
int so_set_stream(so_ctx* ctx, void* cuda_stream);

/* ---- map (LocalMap) ------------------------------------------------------------------------------- */
/* replaces: slam.localMap.lineRes_/planeRes_ = ... (laserMapping.cpp:102-103,648-649). Rebuilds the neighbour
 * index when plane_res changes (cell edge derives from sqrt(3*planeRes)). */
int so_map_set_resolution(so_ctx* ctx, float line_res, float plane_res);
/* replaces: LocalMap::setOrigin (LocalMap.h:146-164): origin_ = -blockOf(t).  out_origin may be NULL. */
int so_map_set_origin(so_ctx* ctx, const double t_w_cur[3], int32_t out_origin[3]);
int so_map_get_origin(so_ctx* ctx, int32_t out_origin[3]);
/* replaces: LocalMap::shiftMap (LocalMap.h:169-287): roll the 21x21x11 block grid so the sensor block stays in
 * [3,17]x[3,17]x[3,7]; blocks rolled off the grid are dropped.  Returns the sensor block in out_ijk. */
int so_map_shift(so_ctx* ctx, const double t_w_cur[3], int32_t out_ijk[3]);
/* Load an ALREADY voxel-filtered world-frame surf cloud as the whole map (no filtering): replay / localization
 * mode with a prebuilt map.  Point ids (so_corr.nn) are indices into this array. */
int so_map_set_points(so_ctx* ctx, const void* xyzi, size_t n, size_t stride_bytes, size_t intensity_offset);
/* Same for the edge clouds (MapBlock::pedge_pc_): an already lineRes-filtered world-frame edge cloud as the whole edge map. */
int so_map_set_edge_points(so_ctx* ctx, const void* xyzi, size_t n, size_t stride_bytes, size_t intensity_offset);
/* replaces: LocalMap::addSurfPointCloud (LocalMap.h:591-645): bin world-frame points into blocks, voxel-centroid
 * filter every touched block at leaf planeRes (pcl::VoxelGrid semantics), rebuild the neighbour index. */
int so_map_add_surf(so_ctx* ctx, const void* xyzi, size_t n, size_t stride_bytes, size_t intensity_offset);
/* replaces: LocalMap::addEdgePointCloud (LocalMap.h:529-589): the same insert at leaf lineRes into the edge clouds. */
int so_map_add_edge(so_ctx* ctx, const void* xyzi, size_t n, size_t stride_bytes, size_t intensity_offset);
/* replaces: LidarSLAM::transformAndAddToMap(cloud, world_cloud, false) (LidarSlam.cpp:60-80): transform the
 * sensor-frame scan by pose (utils::TransformPoint: double math, float store) on the device, then so_map_add_surf. */
int so_map_add_scan(so_ctx* ctx, const void* xyzi, size_t n, size_t stride_bytes, size_t intensity_offset, const double pose[7]);
/* Same insert for the surf cloud most recently passed to so_register, which is still in device memory: the live loop
 * Localization() -> transformAndAddToMap() (LidarSlam.cpp:155-171) without uploading the scan a second time. */
int so_map_add_registered_scan(so_ctx* ctx, const double pose[7]);
/* ... and transformAndAddToMap(cloud, world_cloud, true) for the edge cloud. */
int so_map_add_scan_edge(so_ctx* ctx, const void* xyzi, size_t n, size_t stride_bytes, size_t intensity_offset, const double pose[7]);
/* replaces: LocalMap::get5x5LocalMapFeatureSize (LocalMap.h:291-318) */
int so_map_counts_5x5(so_ctx* ctx, const int32_t ijk[3], int32_t* n_edge, int32_t* n_surf);
/* replaces: LocalMap::getAllLocalMap (mode 0, LocalMap.h:647-658) / get5x5LocalMap(pos) (mode 1, :660-687).
 * Writes packed float4 {x,y,z,intensity} in block order then point-id order; *n_out = points available. */
int so_map_download(so_ctx* ctx, int mode, const int32_t ijk[3], float* out_xyzi, size_t cap_points, size_t* n_out);
size_t so_map_size(so_ctx* ctx);

/* ---- scan pre-filter (the step right before the path) ------------------------------------------ */
/* replaces: laserMapping::adjustVoxelSize (src/LaserMapping/laserMapping.cpp:600-651) for the surf cloud: with
 * auto_voxel_size, pick config_.lineRes/planeRes from the scan statistics (mean|x| * mean|y| * mean|z| < 25 -> 0.1/0.2,
 * > 65 -> 0.4/0.8, else unchanged), then pcl::VoxelGrid(planeRes) on the sensor-frame cloud, then push the resolutions
 * into the map (slam.localMap.lineRes_/planeRes_ = config_, :648-649).  line_res/plane_res are in/out.  Writes up to cap
 * packed float4 {x,y,z,intensity} points ordered by voxel index; *n_out = points produced. */
int so_scan_prefilter(so_ctx* ctx, const void* xyzi, size_t n, size_t stride_bytes, size_t intensity_offset, int auto_voxel_size,
                      float* line_res, float* plane_res, float* out_xyzi, size_t cap_points, size_t* n_out, double* average_distance);

/* replaces: featureExtraction::removePointDistortion<BufferType> (src/FeatureExtraction/featureExtraction.cpp:222-314).
 * points: n records of stride_bytes, float x,y,z first and the per-point float `time` (seconds after lidar_start_time,
 * point_os::PointcloudXYZITR: stride 32, time at byte 20) at time_offset; x,y,z are rewritten in place into the sensor frame
 * at lidar_start_time, non-finite points are left alone.  The pose buffer (MapRingBuffer::measMap_) is passed as n_samples
 * strictly ascending stamps + poses {tx,ty,tz,qx,qy,qz,qw}; interpolation = std::map::upper_bound, slerp + lerp, with the
 * `< 0.0001` rewind of :258-260.  imu_only != 0 is the Imu::Ptr instantiation: samples contribute rotation only and the motion
 * is conjugated by the IMU-lidar extrinsic T_i_l (parameter.cpp:192-193).  start_pose_out (optional) receives
 * {t_w_original_l, q_w_original_l} (:283-289).  *n_past_end (optional) counts points stamped after the last sample -- the
 * reference dereferences end() there and relies on synchronize_measurements (:187-196) to prevent it; the last interval is
 * extrapolated here. */
int so_scan_deskew(so_ctx* ctx, void* points, size_t n, size_t stride_bytes, size_t time_offset, double lidar_start_time,
                   const double* sample_times, const double* sample_poses, size_t n_samples, int imu_only, const double T_i_l[7],
                   double start_pose_out[7], size_t* n_past_end);
/* replaces: featureExtraction::uniformFeatureExtraction (featureExtraction.cpp:504-525): every skip_num-th point from index 1
 * that differs from its predecessor -- |dx| > 1e-7 || |dy| > 1e-7 || (|dz| > 1e-7 && x*x+y*y+z*z > block_range^2), with that
 * precedence -- is emitted as packed float4 {x, y, z, intensity = time}, input order.  int_abs != 0 evaluates the
 * unqualified abs() of :515-517 as ::abs(int) (its meaning when only <cmath> declared abs); 0 = the float overload.
 * *n_out = points produced (may exceed cap_points; only cap_points are written). */
int so_scan_extract_uniform(so_ctx* ctx, const void* points, size_t n, size_t stride_bytes, size_t time_offset, int skip_num,
                            float block_range, int int_abs, float* out_xyzi, size_t cap_points, size_t* n_out);

/* ---- registration (LidarSLAM) --------------------------------------------------------------------- */
/* replaces: LidarSLAM::Localization(true, predictodom, position, edge, planner, t) -> performLocalizationAndMapping
 * (LidarSlam.cpp:30-51,107-171), excluding the map insert at its end (call so_map_add_surf with the
 * transformed scan, as transformAndAddToMap does, LidarSlam.cpp:60-80).
 * edge cloud: processed by the edge / line branch (processEdgeFeatures, LidarSlam.cpp:310-321 -> ComputeLineDistanceParameters
 * :402-493, EdgeAnalyticCostFunction lidarOptimization.cpp:12-47) against the edge map; empty in the shipped pipeline
 * (featureExtraction.cpp:429-436), in which case the branch is idle exactly as upstream (LidarSlam.cpp:311). */
int so_register(so_ctx* ctx,
                const void* surf_xyzi, size_t n_surf, const void* edge_xyzi, size_t n_edge,
                size_t stride_bytes, size_t intensity_offset,
                const double pose_in[7], const so_icp_opts* opts, so_icp_result* out);

/* so_register of the cloud the LAST so_scan_prefilter call produced, which is still on the device (laserMapping.cpp:639-649 filters
 * the scan, :713-714 registers the filtered cloud): no download + second upload between the two steps.  The caller may pass
 * out_xyzi = NULL / cap = 0 to so_scan_prefilter when it does not need the filtered cloud on the host.  Fails with SO_ERR_ARG when
 * anything that reuses the scan buffers ran in between.  so_map_add_registered_scan afterwards inserts that same cloud. */
int so_register_prefiltered(so_ctx* ctx, const double pose_in[7], const so_icp_opts* opts, so_icp_result* out);

/* Batched replay against the frozen map (BASELINE cfg4): n_scans independent Localization() calls.
 * scans: n_scans consecutive clouds, scan s has n_points[s] points starting at point offset sum(n_points[0..s)).
 * poses_in: n_scans x 7.  results: n_scans structs.  n_scans <= so_config.max_batch. */
int so_register_batch(so_ctx* ctx, const void* surf_xyzi, const uint32_t* n_points, size_t n_scans,
                      size_t stride_bytes, size_t intensity_offset,
                      const double* poses_in, const so_icp_opts* opts, so_icp_result* results);
/* Same with an edge cloud per scan (the line branch of so_register, per scan): edge_xyzi holds the n_scans edge clouds back to
 * back, n_edge[s] points each (0 allowed); their total must fit so_config.max_scan_points.  Upstream the edge cloud is empty
 * (featureExtraction.cpp:429-436), so this is the batched form of a dormant branch -- complete, not tuned. */
int so_register_batch_edges(so_ctx* ctx, const void* surf_xyzi, const uint32_t* n_points, const void* edge_xyzi, const uint32_t* n_edge,
                            size_t n_scans, size_t stride_bytes, size_t intensity_offset,
                            const double* poses_in, const so_icp_opts* opts, so_icp_result* results);
/* Same, inputs already resident in device memory: packed float4 points (d_scans_xyzi), host n_points / poses.
 * Used for the device-resident throughput figure; results are still copied back to host structs. */
int so_register_batch_device(so_ctx* ctx, const void* d_scans_xyzi, const uint32_t* n_points, size_t n_scans,
                             const double* poses_in, const so_icp_opts* opts, so_icp_result* results);

/* Sharded-replay plumbing (BASELINE cfg4; SURVEY 8e "K8"): a caller-owned DEVICE buffer of cap_rows x 8 doubles.  Every following
 * so_register / so_register_batch* call appends one row per scan, starting at first_row and advancing: {pose_opt[7] (the
 * optimiser output, so_icp_result.pose_opt), status + 256 * n_iterations}; scans that were not registered (soft statuses) carry
 * their prior.  The rows are written by a kernel on the context stream right after the last optimiser step, so a collective
 * enqueued on that stream (ncclAllGather / torch.distributed.all_gather_into_tensor) gathers the poses of a whole replay without
 * a host round trip.  d_rows == NULL detaches the sink.  The reference has no counterpart (single process, one scan at a time). */
int so_set_pose_sink(so_ctx* ctx, void* d_rows, size_t cap_rows, size_t first_row);

/* Test hook: so_register (surf cloud only) with the neighbour SEARCH replaced by caller-supplied neighbour sets --
 * nn_ids[it][i][0..4] = the five map point ids (so_map_set_points order, ascending distance; 0xFFFFFFFF = no result) that
 * findNearestNeighbors (LidarSlam.cpp:720-747) returned for scan point i in ICP iteration it, n_trace_iters iterations of them.
 * Everything after the search (the 3*planeRes gate on the 5th distance, PCA, plane fit, solve, covariance) runs as in
 * so_register.  Feeding the neighbour sets of the reference's own octree (flann/octree.h:1004-1055, compiled verbatim in
 * oracle/_ref) proves that exactness of the k-NN is the only deviation of this library from the reference path. */
int so_register_injected(so_ctx* ctx, const void* surf_xyzi, size_t n_surf, size_t stride_bytes, size_t intensity_offset,
                         const double pose_in[7], const so_icp_opts* opts, const uint32_t* nn_ids, int32_t n_trace_iters,
                         so_icp_result* out);

/* Stage-level entry points used by the parity tests (same kernels as so_register). */
/* processPlannerFeatures at a fixed pose (LidarSlam.cpp:323-344): fills corr[n], hist_obs[9], hist_rej[7]. */
int so_correspond(so_ctx* ctx, const void* surf_xyzi, size_t n, size_t stride_bytes, size_t intensity_offset,
                  const double pose[7], int32_t max_surface_features, so_corr* corr, int32_t hist_obs[9], int32_t hist_rej[7]);
/* processEdgeFeatures at a fixed pose (LidarSlam.cpp:310-321): fills corr[n], hist_rej_line[7]. */
int so_correspond_edge(so_ctx* ctx, const void* edge_xyzi, size_t n, size_t stride_bytes, size_t intensity_offset, const double pose[7],
                       so_edge_corr* corr, int32_t hist_rej_line[7]);
/* Robustified normal equations at `pose` over the correspondences of the last so_correspond call:
 * H = sum rho' J^T J (row-major 6x6), g = sum rho' J^T r, cost = 1/2 sum rho (lidarOptimization.cpp:55-80 + Ceres corrector). */
int so_evaluate(so_ctx* ctx, const double pose[7], double H[36], double g[6], double* cost);

/* ---- neighbour search (LocalMap::nearestKSearchSurf, LocalMap.h:481-525) --------------------------- */
/* k nearest map points IN THE QUERY'S OWN 50 m BLOCK for nq world-frame queries (packed xyz floats, 12 B stride or
 * given stride).  max_d2 > 0: radius-bounded (neighbours with d2 > max_d2 are not returned); max_d2 <= 0: exact,
 * unbounded.  idx = map point id or 0xFFFFFFFF, d2 = squared distance with the reference's rounding
 * (flann/octree.h:95-102).  Ordering: ascending (d2, id).  1 <= k <= 8.
 * so_knn_device takes device pointers for all three arrays (float4 queries). */
int so_knn(so_ctx* ctx, const float* q_xyz, size_t nq, size_t stride_bytes, int k, float max_d2, uint32_t* idx, float* d2);
int so_knn_device(so_ctx* ctx, const void* d_q_xyzw, size_t nq, int k, float max_d2, uint32_t* d_idx, float* d_d2);

/* ---- feature sampling (calculateSamplingRate + shouldProcessPoint, LidarSlam.cpp:346-359) ---------- */
/* The indices of an n-point scan that a registration capped at max_surface_features processes (ascending; all of 0..n-1 when
 * the cap is <= 0 or >= n).  Host-only (needs no device): this is the list so_register uses to upload the processed points of a
 * capped scan ahead of the rest of the cloud.  Writes min(count, cap) indices to out (may be NULL with cap 0) and the count to
 * *n_out.  Returns SO_OK or SO_ERR_ARG. */
int so_sampling_indices(uint32_t n, int32_t max_surface_features, uint32_t* out, size_t cap, size_t* n_out);

/* ---- instrumentation ------------------------------------------------------------------------------ */
/* Number of this library's kernels launched since the last reset (bench.py gpu_launches). */
uint64_t so_kernel_launches(so_ctx* ctx, int reset);
/* Bytes this library copied host->device / device->host for scans, poses, optimiser state and results since the
 * last reset (bench.py e2e accounting). */
int so_bytes_copied(so_ctx* ctx, uint64_t* h2d, uint64_t* d2h, int reset);
/* Accumulated device time (ms, CUDA events on the context stream) and launch count of a kernel class since the last
 * reset; classes: 0 k_knn_scan (scan k-NN; k_knn_fit = k-NN + plane fit in a SO_BUILD_FUSED_MATCH build), 1 k_evaluate +
 * k_lm_step (LM step), 2 k_knn (so_knn*), 3 scan ordering / map build / map insert / scan preparation, 4 first evaluation of
 * a solve (k_evaluate + k_lm_step iteration zero), 5 k_fit (plane fit; empty in a fused build).
 * Profiling mode serialises kernels; enable only for roofline runs. */
/* Build-time switches of the loaded library: bit 0 (SO_BUILD_FUSED_MATCH) = search and fit run as one kernel. */
#define SO_BUILD_FUSED_MATCH 1
int so_build_flags(void);
int so_profile_enable(so_ctx* ctx, int on);
int so_profile_get(so_ctx* ctx, int kernel_class, double* ms, uint64_t* launches, int reset);

#ifdef __cplusplus
}
#endif
#endif /* SUPERODOM_B200_H */
