// so_ctx.cuh -- the context behind the C ABI: owns every device buffer, the stream, CUDA graphs and host mirrors.
#pragma once
#include <string>
#include <vector>

#include "so_icp.cuh"

namespace so {

#define SO_CUDA_TRY(expr)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            so::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));                   \
            return SO_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

void set_error(const std::string& s);

struct ProfileSlot { double ms = 0.0; uint64_t launches = 0; };

// One per-block cloud family of the LocalMap (surf or edge): points in id order, the sorted hash grid, block tables.
struct MapStore {
    float res = 0.4f;                              // planeRes_ (surf) / lineRes_ (edge): voxel leaf and gate / search radius
    uint32_t n = 0;                                // points held (on-grid)
    float4* d_xyzi = nullptr;                      // [max_map] id order, w = intensity
    float4* d_sorted = nullptr;                    // [max_map] sorted by (slot, cell), w = bitcast(id)
    int32_t* d_block_slot = nullptr;               // [4851]
    int32_t* d_block_count = nullptr;              // [4851]
    uint32_t* d_cell_start = nullptr;              // [cell_cap]
    size_t cell_cap = 0;
    std::vector<int32_t> h_block_count, h_block_slot;   // host mirrors
    int n_slots = 0;
    int nb = 64;                                   // cells per block axis
    bool dirty = true;
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    so_config cfg{};

    // ---- maps: raw (id order) + sorted hash grid; one store for surf points, one for edge points -------------------
    int32_t origin[3] = {kW / 2, kH / 2, kD / 2};  // LocalMap() ctor (LocalMap.h:141-144)
    MapStore surf, edge;                           // psurf_pc_ / pedge_pc_ of every MapBlock
    // scratch shared by both stores (builds are sequential on the context stream)
    uint64_t* d_keys = nullptr;                    // [max_map] sort keys (in)
    uint64_t* d_keys_out = nullptr;                // [max_map]
    uint32_t* d_vals = nullptr;                    // [max_map] ids (in)
    uint32_t* d_vals_out = nullptr;                // [max_map]
    int32_t* d_block_of_point = nullptr;           // [max_map] grid block of each raw point or -1
    void* d_cub_tmp = nullptr;
    size_t cub_tmp_bytes = 0;
    void* d_insert_info = nullptr;                 // device-side bookkeeping of a map insert (InsertInfo + touched-block flags)
    void* h_insert_info = nullptr;                 // pinned read-back target

    // ---- scans / correspondences / optimiser state -------------------------------------------------------------
    uint32_t max_batch = 1;
    size_t scan_cap = 0;                           // points over the whole batch
    size_t prefiltered_n = SIZE_MAX;               // points so_scan_prefilter left in d_scan_sorted (SIZE_MAX: none / overwritten since)
    uint32_t last_scan_n = 0;                      // points of the scan so_register uploaded last (d_scan[0..n), original order)
    float4* d_scan = nullptr;                      // upload target for host scans (original order)
    float4* d_scan_sorted = nullptr;               // cell-ordered copy the kernels read; w = original index
    uint64_t* d_skeys = nullptr; uint64_t* d_skeys_out = nullptr;   // scan sort keys (scan id << 32 | cell)
    uint32_t* d_svals = nullptr; uint32_t* d_svals_out = nullptr;
    void* d_sort_tmp = nullptr; size_t sort_tmp_bytes = 0;
    void* d_sort_tmp2 = nullptr;                   // second radix-sort scratch: scan ordering of the chunks that run on aux_stream
    NnBuf nn{};
    uint32_t* d_offset = nullptr;
    IcpState* d_state = nullptr;
    IcpState* h_state = nullptr;                   // pinned
    uint32_t* h_offset = nullptr;                  // pinned
    uint32_t* h_eoffset = nullptr;                 // second half of the same allocation
    bool eoffset_dirty = false;                    // d_eoffset holds non-zero entries from the last batch
    double* d_partials = nullptr;
    uint32_t* d_counters = nullptr;
    int32_t* d_hist = nullptr;
    CorrBuf corr{};
    // edge / line branch (single-scan so_register only)
    size_t edge_cap = 0; uint32_t edge_grid_cap = 0;
    float4* d_escan = nullptr; uint32_t* d_eoffset = nullptr; EdgeBuf ebuf{};
    uint32_t grid_x_cap = 0;
    double* d_pose_sink = nullptr; size_t sink_cap = 0, sink_cursor = 0;   // so_set_pose_sink: device rows the batch calls append to
    uint32_t* d_inject = nullptr; size_t inject_cap = 0;   // so_register_injected: caller-supplied neighbour ids [iters][n][5]
    void* h_stage = nullptr;                       // pinned staging for strided host clouds
    size_t h_stage_bytes = 0;
    // so_register of a DECIMATED scan (max_surface_features < n): only the points shouldProcessPoint keeps go up before the
    // registration is launched; the whole cloud (what so_map_add_registered_scan inserts) follows on the copy stream while it runs
    float4* d_dec = nullptr;                       // [kPrepareSmallCap] the kept points, index order, w = intensity
    std::vector<uint32_t> dec_idx;                 // kept indices for (dec_n, dec_mf): the list depends on nothing else
    uint32_t dec_n = 0; int32_t dec_mf = 0;
    cudaEvent_t ev_dec = nullptr;                  // the deferred upload of the whole cloud has landed in d_scan
    struct { const void* src = nullptr; size_t n = 0, stride = 0, ioff = 0, stage_off = 0; } defer;      // pending whole-cloud upload (src == nullptr: none)
    bool no_dec_upload = false;                    // SO_NO_DEC_UPLOAD: A/B aid, always upload the whole cloud first

    // ---- k-NN scratch -------------------------------------------------------------------------------------------
    float4* d_q = nullptr; uint32_t* d_knn_idx = nullptr; float* d_knn_d2 = nullptr; size_t knn_cap = 0;
    uint32_t* d_qkeys = nullptr; uint32_t* d_qkeys_out = nullptr; uint32_t* d_qvals = nullptr; uint32_t* d_qvals_out = nullptr;
    void* d_qsort_tmp = nullptr; size_t qsort_tmp_bytes = 0; size_t qsort_cap = 0;      // cell ordering of so_knn* queries

    // ---- CUDA graph cache for the ICP schedule (one entry per batch chunk shape) ------------------------------------
    // Everything a captured schedule bakes into its kernel arguments: the chunk shape and both MapViews (pointers, origin, grid,
    // resolutions).  Keyed on these VALUES, not on a "map changed" counter, so that a map insert that leaves the index arrays
    // where they are (the usual live-SLAM case) keeps hitting the cache.
    struct GraphKey {
        uint32_t first, count, grid_x, grid_e; int32_t iters, lm;
        const void* ptrs[8]; int32_t origin[3]; int32_t nb[2]; float res[2];
    };
    struct GraphSlot {
        cudaGraphExec_t exec = nullptr;
        GraphKey key{}; bool is_loop = false; uint64_t used = 0;
    };
    GraphSlot graphs[20];
    uint64_t graph_clock = 0;
    uint64_t map_epoch = 1;
    bool force_key64 = false;                      // SO_FORCE_KEY64: test aid, take the 64-bit scan-order key path even when 32 bits suffice
    bool no_coop_knn = false;                      // SO_NO_COOP_KNN: A/B aid, always one query per thread
    bool no_small_prepare = false;                 // SO_NO_SMALL_PREPARE: A/B aid, always the device-wide scan ordering
    bool no_fused_lm = false;                      // SO_NO_FUSED_LM: A/B aid, always the two-kernel evaluation + optimiser step
    bool single_stream = false;                    // SO_SINGLE_STREAM: tuning aid, all chunks on `stream`
    int chunk_override = 0;                        // SO_CHUNKS: tuning aid, upload/compute chunks per host batch (0 = built-in rule)
    bool no_cond_graph = false;                    // conditional nodes unavailable (or SO_NO_COND_GRAPH): unrolled schedule
    cudaStream_t aux_stream = nullptr;             // odd chunks of a batch run here, concurrently with the even ones on `stream`:
                                                   // one chunk's serial optimiser steps and kernel tails hide under the other's wide kernels
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t copy_stream = nullptr;            // H2D of batch chunks overlaps the previous chunk's kernels
    cudaEvent_t ev_copy[17] = {};           // [0..15]: per upload chunk; [16]: guard

    // ---- instrumentation ----------------------------------------------------------------------------------------
    uint64_t launches = 0;
    uint64_t bytes_h2d = 0, bytes_d2h = 0;         // host<->device bytes moved by registration / k-NN calls
    bool profiling = false;
    ProfileSlot prof[6];
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp0 = nullptr, evp1 = nullptr;
};

// so_map.cu
int map_alloc(Ctx* c);
void map_free(Ctx* c);
int map_add_points(Ctx* c, MapStore& ms, uint32_t n_new);   // ms.d_xyzi[n .. n+n_new) holds the new points: voxel-filter touched blocks, rebuild
int map_transform_tail(Ctx* c, MapStore& ms, uint32_t n_new, const double pose[7]);   // sensor-frame tail points -> world frame
int scan_voxel_filter(Ctx* c, uint32_t n, float leaf, uint32_t* n_out);
// so_scan.cu
struct DeskewParams {
    double start_time;                     // lidar_start_time
    const double* times;                   // n_samples ascending sample stamps (device)
    const double* poses;                   // n_samples x {tx,ty,tz,qx,qy,qz,qw} (device; translation zeroed for IMU samples)
    uint32_t n_samples, n_smem;            // n_smem: stamps cached in shared memory
    int imu_only;
    double q0_conj[4], p0[3];              // T_w_original: conjugate of its (normalised) rotation, its position
    double q_il[4], t_il[3], q_li[4], t_li[3];   // T_i_l and its inverse (parameter.cpp:192-193)
    uint32_t* past_end;                    // points later than the last sample
};
void launch_deskew(float4* pts, uint32_t n, const DeskewParams& P, cudaStream_t st);
int scan_extract_uniform(Ctx* c, uint32_t n, uint32_t skip, float block_range, int int_abs, uint32_t* n_out);   // d_scan -> d_scan_sorted   // d_scan -> d_scan_sorted (VoxelGrid on a scan)
int map_rebuild(Ctx* c, MapStore& ms);    // (re)bin, drop off-grid points, sort, build cell table
MapView map_view(const Ctx* c, const MapStore& ms);
int map_cells_per_block(float plane_res);
int scan_sort_alloc(Ctx* c);             // temp storage for the per-registration scan sort
int query_sort(Ctx* c, size_t n);         // d_qkeys/d_qvals -> *_out (allocates on growth)
int query_sort_reserve(Ctx* c, size_t n);
int scan_sort(Ctx* c, size_t first, size_t n, int n_scans, int cell_bits, bool key32, cudaStream_t st);   // d_skeys/d_svals[first..first+n) -> *_out

}  // namespace so
