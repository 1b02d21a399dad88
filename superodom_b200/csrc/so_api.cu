// so_api.cu -- the C ABI (include/superodom_b200.h): context, uploads, the CUDA-graph ICP schedule, results.
// Host code above the kernels stays C++ (the reference's host language); no torch types anywhere.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "so_ctx.cuh"
#include "so_chunks.h"
#include "so_knn.cuh"

namespace so {

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }

static int fail(int code, const std::string& msg) { set_error(msg); return code; }

// ---------------------------------------------------------------------------------------------- host pose helpers
// tf2 Matrix3x3(q).getRPY + Quaternion::setRPY round trip: LidarSLAM::MannualYawCorrection (LidarSlam.cpp:891-913)
static void manual_yaw_correction(const double last[7], double T[7], double yaw_ratio) {
    double tn, rn;
    rel_motion(last, T, &tn, &rn);
    const float translation_norm = float(tn);
    const double x = T[3], y = T[4], z = T[5], w = T[6];
    const double d = x * x + y * y + z * z + w * w, s = 2.0 / d;
    const double xs = x * s, ys = y * s, zs = z * s, wx = w * xs, wy = w * ys, wz = w * zs;
    const double xx = x * xs, xy = x * ys, xz = x * zs, yy = y * ys, yz = y * zs, zz = z * zs;
    const double m00 = 1.0 - (yy + zz), m01 = xy - wz, m02 = xz + wy, m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
    double roll, pitch, yaw;
    if (std::fabs(m20) >= 1) {
        yaw = 0;
        if (m20 < 0) { pitch = M_PI / 2.0; roll = std::atan2(m01, m02); }
        else { pitch = -M_PI / 2.0; roll = std::atan2(-m01, -m02); }
    } else {
        pitch = -std::asin(m20);
        roll = std::atan2(m21 / std::cos(pitch), m22 / std::cos(pitch));
        yaw = std::atan2(m10 / std::cos(pitch), m00 / std::cos(pitch));
    }
    const double cyaw = yaw + double(translation_norm) * yaw_ratio * M_PI / 180;
    const double hy = cyaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
    const double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
    double q[4] = {sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy};
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) T[3 + i] = q[i] / n;
}

static inline int block_coord_h(double v, int origin) { int c = int(v / kBlock); if (v < 0) c--; return c + origin; }

// LocalMap::get5x5LocalMapFeatureSize (LocalMap.h:291-318) on the host mirror of the block counts
static int counts_5x5(const Ctx* c, const MapStore& ms, const int32_t ijk[3]) {
    (void)c;
    int n = 0;
    for (int i = ijk[0] - 2; i <= ijk[0] + 2; ++i)
        for (int j = ijk[1] - 2; j <= ijk[1] + 2; ++j)
            for (int k = ijk[2] - 1; k <= ijk[2] + 1; ++k)
                if (i >= 0 && i < kW && j >= 0 && j < kH && k >= 0 && k < kD) n += ms.h_block_count[i + kW * j + kW * kH * k];
    return n;
}

static int ensure_maps(Ctx* c) {
    if (c->surf.dirty) { int rc = map_rebuild(c, c->surf); if (rc) return rc; }
    if (c->edge.dirty) { int rc = map_rebuild(c, c->edge); if (rc) return rc; }
    return SO_OK;
}

// ---------------------------------------------------------------------------------------------- allocation
static int ctx_alloc(Ctx* c) {
    SO_CUDA_TRY(cudaSetDevice(c->device));
    SO_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SO_CUDA_TRY(cudaEventCreate(&c->ev0)); SO_CUDA_TRY(cudaEventCreate(&c->ev1));
    SO_CUDA_TRY(cudaEventCreate(&c->evp0)); SO_CUDA_TRY(cudaEventCreate(&c->evp1));
    SO_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    SO_CUDA_TRY(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    SO_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    SO_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (auto& e : c->ev_copy) SO_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int rc = map_alloc(c);
    if (rc) return rc;
    c->max_batch = c->cfg.max_batch;
    c->scan_cap = size_t(c->cfg.max_scan_points) * c->max_batch;
    c->grid_x_cap = (c->cfg.max_scan_points + kThreads - 1) / kThreads;
    SO_CUDA_TRY(cudaMalloc(&c->d_scan, c->scan_cap * sizeof(float4)));
    SO_CUDA_TRY(cudaMalloc(&c->d_dec, size_t(kPrepareSmallCap) * sizeof(float4)));
    SO_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_dec, cudaEventDisableTiming));
    SO_CUDA_TRY(cudaMalloc(&c->d_scan_sorted, c->scan_cap * sizeof(float4)));
    SO_CUDA_TRY(cudaMalloc(&c->d_skeys, c->scan_cap * sizeof(uint64_t) + 64));      // + slack: scan_voxel_filter lays [count 16 B][heads][ranks] over it
    SO_CUDA_TRY(cudaMalloc(&c->d_skeys_out, c->scan_cap * sizeof(uint64_t)));
    SO_CUDA_TRY(cudaMalloc(&c->d_svals, c->scan_cap * sizeof(uint32_t)));
    SO_CUDA_TRY(cudaMalloc(&c->d_svals_out, c->scan_cap * sizeof(uint32_t)));
    SO_CUDA_TRY(cudaMalloc(&c->nn.pos, c->scan_cap * 5 * sizeof(uint32_t)));
    SO_CUDA_TRY(cudaMalloc(&c->nn.pts, c->scan_cap * 5 * sizeof(float4)));
    SO_CUDA_TRY(cudaMalloc(&c->nn.d5, c->scan_cap * sizeof(float)));
    SO_CUDA_TRY(cudaMalloc(&c->nn.pre, c->scan_cap));
    SO_CUDA_TRY(cudaMalloc(&c->nn.vq, c->scan_cap * sizeof(float4)));
    c->nn.cap = c->scan_cap;
    { int rc2 = scan_sort_alloc(c); if (rc2) return rc2; }
    SO_CUDA_TRY(cudaMalloc(&c->d_offset, c->max_batch * sizeof(uint32_t)));
    SO_CUDA_TRY(cudaMalloc(&c->d_state, c->max_batch * sizeof(IcpState)));
    SO_CUDA_TRY(cudaMallocHost(&c->h_state, c->max_batch * sizeof(IcpState)));
    SO_CUDA_TRY(cudaMallocHost(&c->h_offset, 2 * c->max_batch * sizeof(uint32_t)));      // [0, max_batch): scan offsets; then the edge clouds' offsets
    c->h_eoffset = c->h_offset + c->max_batch;
    c->edge_cap = c->cfg.max_scan_points;
    c->edge_grid_cap = uint32_t((c->edge_cap + kThreads - 1) / kThreads);
    SO_CUDA_TRY(cudaMalloc(&c->d_partials, size_t(c->max_batch) * (c->grid_x_cap + c->edge_grid_cap) * kAcc * sizeof(double)));
    SO_CUDA_TRY(cudaMalloc(&c->d_escan, c->edge_cap * sizeof(float4)));
    SO_CUDA_TRY(cudaMalloc(&c->d_eoffset, c->max_batch * sizeof(uint32_t)));
    SO_CUDA_TRY(cudaMemset(c->d_eoffset, 0, c->max_batch * sizeof(uint32_t)));
    SO_CUDA_TRY(cudaMalloc(&c->ebuf.a, c->edge_cap * sizeof(double4)));
    SO_CUDA_TRY(cudaMalloc(&c->ebuf.b, c->edge_cap * sizeof(double4)));
    SO_CUDA_TRY(cudaMalloc(&c->ebuf.flags, c->edge_cap * sizeof(uchar4)));
    SO_CUDA_TRY(cudaMalloc(&c->ebuf.nn, c->edge_cap * 10 * sizeof(uint32_t)));
    SO_CUDA_TRY(cudaMalloc(&c->ebuf.selmask, c->edge_cap * sizeof(uint32_t)));
    c->ebuf.scan = c->d_escan; c->ebuf.offset = c->d_eoffset;
    SO_CUDA_TRY(cudaMalloc(&c->d_counters, c->max_batch * sizeof(uint32_t)));
    SO_CUDA_TRY(cudaMalloc(&c->d_hist, c->max_batch * kHistStride * sizeof(int32_t)));
    SO_CUDA_TRY(cudaMemset(c->d_counters, 0, c->max_batch * sizeof(uint32_t)));
    SO_CUDA_TRY(cudaMemset(c->d_hist, 0, c->max_batch * kHistStride * sizeof(int32_t)));
    SO_CUDA_TRY(cudaMalloc(&c->corr.nd, c->scan_cap * sizeof(double4)));
    SO_CUDA_TRY(cudaMalloc(&c->corr.w, c->scan_cap * sizeof(double)));
    SO_CUDA_TRY(cudaMalloc(&c->corr.flags, c->scan_cap * sizeof(uchar4)));
    SO_CUDA_TRY(cudaMalloc(&c->corr.nn, size_t(c->cfg.max_scan_points) * 5 * sizeof(uint32_t)));
    SO_CUDA_TRY(cudaMalloc(&c->corr.nn_d2, size_t(c->cfg.max_scan_points) * 5 * sizeof(float)));
    return SO_OK;
}

static void ctx_free(Ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& g : c->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (auto& e : c->ev_copy) if (e) cudaEventDestroy(e);
    map_free(c);
    cudaFree(c->d_scan_sorted); cudaFree(c->d_skeys); cudaFree(c->d_skeys_out); cudaFree(c->d_svals); cudaFree(c->d_svals_out);
    cudaFree(c->d_sort_tmp); cudaFree(c->d_sort_tmp2); cudaFree(c->nn.pos); cudaFree(c->nn.pts); cudaFree(c->nn.d5); cudaFree(c->nn.pre); cudaFree(c->nn.vq);
    cudaFree(c->d_scan); cudaFree(c->d_offset); cudaFree(c->d_state); cudaFreeHost(c->h_state); cudaFreeHost(c->h_offset);
    cudaFree(c->d_partials); cudaFree(c->d_counters); cudaFree(c->d_hist);
    cudaFree(c->d_escan); cudaFree(c->d_eoffset); cudaFree(c->ebuf.a); cudaFree(c->ebuf.b); cudaFree(c->ebuf.flags); cudaFree(c->ebuf.nn); cudaFree(c->ebuf.selmask);
    cudaFree(c->corr.nd); cudaFree(c->corr.w); cudaFree(c->corr.flags); cudaFree(c->corr.nn); cudaFree(c->corr.nn_d2);
    cudaFree(c->d_q); cudaFree(c->d_knn_idx); cudaFree(c->d_knn_d2); cudaFree(c->d_inject);
    cudaFree(c->d_qkeys); cudaFree(c->d_qkeys_out); cudaFree(c->d_qvals); cudaFree(c->d_qvals_out); cudaFree(c->d_qsort_tmp);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    cudaFree(c->d_dec);
    if (c->ev_dec) cudaEventDestroy(c->ev_dec);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evp0) cudaEventDestroy(c->evp0);
    if (c->evp1) cudaEventDestroy(c->evp1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int ensure_stage(Ctx* c, size_t bytes) {
    if (bytes <= c->h_stage_bytes) return SO_OK;
    if (c->h_stage) { cudaDeviceSynchronize(); cudaFreeHost(c->h_stage); }      // no copy out of the old buffer may still be in flight
    c->h_stage = nullptr; c->h_stage_bytes = 0;
    SO_CUDA_TRY(cudaMallocHost(&c->h_stage, bytes));
    c->h_stage_bytes = bytes;
    return SO_OK;
}

// Copy a strided host cloud into device float4 {x,y,z,intensity}.  Packed float4 input goes straight through.
static int upload_cloud(Ctx* c, const void* src, size_t n, size_t stride, size_t ioff, float4* dst) {
    if (n == 0) return SO_OK;
    if (stride == 16 && ioff == 12) {
        SO_CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
        c->bytes_h2d += n * sizeof(float4);
        return SO_OK;
    }
    int rc = ensure_stage(c, n * sizeof(float4));
    if (rc) return rc;
    SO_CUDA_TRY(cudaStreamSynchronize(c->stream));     // staging buffer reuse
    float4* h = static_cast<float4*>(c->h_stage);
    const unsigned char* p = static_cast<const unsigned char*>(src);
    for (size_t i = 0; i < n; ++i, p += stride) {
        float xyz[3], it = 0.f;
        std::memcpy(xyz, p, 12);
        if (ioff + 4 <= stride) std::memcpy(&it, p + ioff, 4);
        h[i] = make_float4(xyz[0], xyz[1], xyz[2], it);
    }
    SO_CUDA_TRY(cudaMemcpyAsync(dst, h, n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    c->bytes_h2d += n * sizeof(float4);
    return SO_OK;
}

// shouldProcessPoint (LidarSlam.cpp:353-359) on the host: the same three IEEE operations as so_icp.cu's should_process (no FMA)
static inline bool should_process_h(uint32_t i, double rate) {
    const double x = double(i) * rate;
    const double rem = x - double(int64_t(x));            // x >= 0 and < 2^32: truncation == floor
    return !(rem + 0.001 > rate);
}

// Indices shouldProcessPoint keeps of an n-point scan capped at max_features (ascending).  A function of (n, max_features) only,
// so the list of the last call is kept.
static const std::vector<uint32_t>& decimation_list(Ctx* c, uint32_t n, int32_t max_features) {
    if (c->dec_n == n && c->dec_mf == max_features && !c->dec_idx.empty()) return c->dec_idx;
    const double rate = 1.0 * max_features / double(n);       // calculateSamplingRate (LidarSlam.cpp:346-351)
    c->dec_idx.clear();
    for (uint32_t i = 0; i < n; ++i) if (should_process_h(i, rate)) c->dec_idx.push_back(i);
    c->dec_n = n; c->dec_mf = max_features;
    return c->dec_idx;
}

static inline float4 read_point(const unsigned char* p, size_t stride, size_t ioff) {
    float xyz[3], it = 0.f;
    std::memcpy(xyz, p, 12);
    if (ioff + 4 <= stride) std::memcpy(&it, p + ioff, 4);
    return make_float4(xyz[0], xyz[1], xyz[2], it);
}

// The pending whole-cloud upload of a decimated so_register (Ctx::defer), on the COPY stream: it only has to land before the scan
// is inserted, so it runs beside the registration instead of in front of it.  No-op when nothing is pending.
static int run_deferred_upload(Ctx* c) {
    if (!c->defer.src) return SO_OK;
    const void* src = c->defer.src;
    c->defer.src = nullptr;
    const size_t n = c->defer.n;
    if (c->defer.stride == 16 && c->defer.ioff == 12) {
        SO_CUDA_TRY(cudaMemcpyAsync(c->d_scan, src, n * sizeof(float4), cudaMemcpyHostToDevice, c->copy_stream));
    } else {
        float4* h = static_cast<float4*>(c->h_stage) + c->defer.stage_off;
        const unsigned char* p = static_cast<const unsigned char*>(src);
        for (size_t i = 0; i < n; ++i, p += c->defer.stride) h[i] = read_point(p, c->defer.stride, c->defer.ioff);
        SO_CUDA_TRY(cudaMemcpyAsync(c->d_scan, h, n * sizeof(float4), cudaMemcpyHostToDevice, c->copy_stream));
    }
    c->bytes_h2d += n * sizeof(float4);
    SO_CUDA_TRY(cudaEventRecord(c->ev_dec, c->copy_stream));
    return SO_OK;
}

// ---------------------------------------------------------------------------------------------- ICP schedule
static BatchView batch_view(const Ctx* c, const float4* scan, uint32_t first = 0) {
    BatchView bv;
    bv.scan = scan; bv.offset = c->d_offset + first; bv.st = c->d_state + first;
    bv.partial_stride = c->grid_x_cap + c->edge_grid_cap; bv.edge_partial_offset = c->grid_x_cap;
    bv.partials = c->d_partials + size_t(first) * bv.partial_stride * kAcc; bv.hist = c->d_hist + size_t(first) * kHistStride;
    const double a = double(std::sqrt(3 * c->surf.res));     // float sqrt of a float product (LidarSlam.cpp:271)
    bv.tukey_a2 = a * a;
    const double al = double(std::sqrt(3 * c->edge.res));     // TukeyLoss(std::sqrt(3*lineRes_)) (LidarSlam.cpp:263)
    bv.tukey_a2_line = al * al;
    bv.counters = nullptr;
    return bv;
}

static inline void count_h2d(Ctx* c, size_t b) { c->bytes_h2d += b; }
static inline void count_d2h(Ctx* c, size_t b) { c->bytes_d2h += b; }

static void timed_launch_begin(Ctx* c) { if (c->profiling) cudaEventRecord(c->evp0, c->stream); }
static void timed_launch_end(Ctx* c, int cls) {
    c->launches++;
    if (!c->profiling) return;
    cudaEventRecord(c->evp1, c->stream);
    cudaEventSynchronize(c->evp1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->evp0, c->evp1);
    c->prof[cls].ms += ms; c->prof[cls].launches++;
}

// A contiguous run of scans of the batch that is uploaded, ordered and registered together.
// grid_in: CTAs covering the longest scan as uploaded (scan ordering); grid_x: CTAs covering the longest scan once it is reduced to the
// points shouldProcessPoint keeps (the ICP kernels); small: every scan's survivors fit k_prepare_small
struct Chunk { uint32_t first, count, pt_first, pt_count, grid_x, grid_e = 0, grid_in = 0; bool small = false; uint32_t max_kept = 0; };

// Once per registration: order every scan by map cell at its prior pose (k_scan_keys -> radix sort -> k_scan_gather, or the one-CTA
// k_prepare_small for small registrations).  compact: drop the points the decimation skips here, once, instead of in every kernel.
static int prepare_scans(Ctx* c, const float4* d_scan_in, const Chunk& ch, cudaStream_t st, bool compact) {
    c->prefiltered_n = SIZE_MAX;                                   // d_scan_sorted is about to be overwritten
    const MapView mv = map_view(c, c->surf);
    const BatchView bv = batch_view(c, d_scan_in, ch.first);
    timed_launch_begin(c);
    // key layout: [scan inside the chunk | dropped (compact only) | cell]; 32-bit keys whenever everything fits
    const uint64_t n_cells = scan_key_space(c->surf.n_slots, mv.nb);       // cell / brick-order keys (so_knn.cuh)
    int cell_bits = 1, scan_bits = 0;
    while (cell_bits < 32 && (uint64_t(1) << cell_bits) <= n_cells) ++cell_bits;       // cells 0..n_cells-1 < mask = 2^cell_bits - 1
    if (compact && ch.small && !c->no_small_prepare) {
        launch_prepare_small(mv, bv, c->d_scan_sorted, ch.count, cell_bits, ch.max_kept, st);
        c->launches++;
        timed_launch_end(c, 3);
        SO_CUDA_TRY(cudaGetLastError());
        return SO_OK;
    }
    while ((1u << scan_bits) < ch.count) ++scan_bits;
    const int extra = compact ? 1 : 0;
    const bool key32 = !c->force_key64 && cell_bits < 32 && cell_bits + extra + scan_bits <= 32;
    if (!key32) cell_bits = 32;
    const uint32_t grid_in = ch.grid_in ? ch.grid_in : ch.grid_x;
    launch_scan_keys(mv, bv, c->d_skeys, c->d_svals, grid_in, ch.count, cell_bits, key32, compact, st);
    int rc = scan_sort(c, ch.pt_first, ch.pt_count, int(ch.count), cell_bits + extra, key32, st);
    if (rc) return rc;
    launch_scan_gather(d_scan_in, c->d_svals_out, c->d_skeys_out, ch.pt_first, c->d_offset + ch.first, ch.pt_count,
                       c->d_scan_sorted + ch.pt_first, cell_bits + extra, key32, st);
    if (compact) { launch_scan_finish(bv, c->d_skeys_out, ch.pt_first, ch.pt_first, cell_bits, key32, ch.count, st); c->launches++; }
    c->launches++;
    timed_launch_end(c, 3);
    SO_CUDA_TRY(cudaGetLastError());
    return SO_OK;
}

// [k_knn_scan, k_fit, k_evaluate<PH_CORR>, k_lm_step, (k_evaluate<PH_EVAL>, k_lm_step) x lm] per ICP iteration; every kernel exits at once when its
// scan is not in the matching phase, so a fixed schedule follows whatever path the device-side state machine takes.
// Preferred form: a CUDA-graph WHILE node around ONE iteration (k_loop_cond ends the loop when every scan of the chunk
// is done); fallback: the schedule unrolled max_icp_iters times.  Returns whether the loop form ran.
static int run_schedule(Ctx* c, const Chunk& ch, int iters, int lm, bool with_nn, bool* was_loop, cudaStream_t run_stream = nullptr,
                        const uint32_t* d_inject = nullptr, int inject_iters = 0) {
    if (!run_stream) run_stream = c->stream;
    const float4* d_scan = c->d_scan_sorted;
    const MapView mv = map_view(c, c->surf);
    BatchView bv = batch_view(c, d_scan, ch.first);
    CorrBuf cb = c->corr;
    if (!with_nn) { cb.nn = nullptr; cb.nn_d2 = nullptr; }
    const uint32_t grid_x = ch.grid_x, n_scans = ch.count, grid_e = ch.grid_e;
    // one or two scans in flight: launch-latency bound -> optimiser step folded into the evaluation kernels (k_evaluate_lm)
    if (n_scans <= 2 && !c->profiling && !c->no_fused_lm) bv.counters = c->d_counters + ch.first;
    bv.coop_knn = n_scans <= 2 && grid_x <= kCoopMaxGrid && !c->no_coop_knn;      // latency-bound search: a warp per query
    const MapView me = map_view(c, c->edge);
    EdgeBuf eb = c->ebuf;
    eb.offset = c->d_eoffset + ch.first;                  // like bv.offset: indexed by the scan's position inside the chunk
    if (!with_nn) { eb.nn = nullptr; eb.selmask = nullptr; }
    *was_loop = false;
    if (d_inject && kCorrLaunches != 2) return fail(SO_ERR_ARG, "so_register_injected needs the split k_knn_scan / k_fit build");
    if (c->profiling || with_nn || d_inject) {
        // profiling mode: one launch at a time, timed with events, and only launches that have work (the host peeks
        // at the phases) so that the per-class average is the duration of a kernel that actually ran
        std::vector<IcpState> peek(n_scans);
        auto any_in = [&](int ph) -> bool {
            cudaMemcpyAsync(peek.data(), c->d_state + ch.first, n_scans * sizeof(IcpState), cudaMemcpyDeviceToHost, c->stream);
            cudaStreamSynchronize(c->stream);
            for (uint32_t s = 0; s < n_scans; ++s) if (peek[s].phase == ph) return true;
            return false;
        };
        for (int it = 0; it < iters; ++it) {
            if (any_in(PH_CORR)) {
                if (kCorrLaunches == 1) { timed_launch_begin(c); launch_match(mv, bv, cb, c->nn, grid_x, n_scans, c->stream); timed_launch_end(c, 0); }
                else {
                    timed_launch_begin(c);
                    if (d_inject) launch_inject(mv, c->surf.d_xyzi, c->surf.n, bv, c->nn, d_inject, inject_iters, grid_x, n_scans, c->stream);
                    else launch_match(mv, bv, cb, c->nn, grid_x, n_scans, c->stream, 1);                                               // k_knn_scan
                    timed_launch_end(c, 0);
                    timed_launch_begin(c); launch_match(mv, bv, cb, c->nn, grid_x, n_scans, c->stream, 2); timed_launch_end(c, 5);      // k_fit
                }
                timed_launch_begin(c); launch_first_eval(bv, cb, grid_x, n_scans, c->stream, &me, &eb, grid_e); c->launches += 1 + (grid_e ? 1 : 0); timed_launch_end(c, 4);
            }
            for (int k = 0; k < lm; ++k)
                if (any_in(PH_EVAL)) { timed_launch_begin(c); launch_evaluate(bv, cb, grid_x, n_scans, c->stream, &eb, grid_e); c->launches += 1 + (grid_e ? 1 : 0); timed_launch_end(c, 1); }
        }
        SO_CUDA_TRY(cudaGetLastError());
        return SO_OK;
    }
    Ctx::GraphKey key;
    std::memset(&key, 0, sizeof(key));                      // padding bytes take part in the memcmp below
    key.first = ch.first; key.count = n_scans; key.grid_x = grid_x; key.grid_e = grid_e; key.iters = iters; key.lm = lm;
    key.ptrs[0] = mv.pts; key.ptrs[1] = mv.cell_start; key.ptrs[2] = mv.block_slot; key.ptrs[3] = mv.block_count;
    key.ptrs[4] = me.pts; key.ptrs[5] = me.cell_start; key.ptrs[6] = me.block_slot; key.ptrs[7] = me.block_count;
    for (int a = 0; a < 3; ++a) key.origin[a] = mv.origin[a];
    key.nb[0] = mv.nb; key.nb[1] = me.nb; key.res[0] = mv.plane_res; key.res[1] = me.plane_res;
    Ctx::GraphSlot* slot = nullptr;
    for (auto& g : c->graphs)
        if (g.exec && std::memcmp(&g.key, &key, sizeof(key)) == 0) slot = &g;
    if (!slot) {
        slot = &c->graphs[0];
        for (auto& g : c->graphs) { if (!g.exec) { slot = &g; break; } if (g.used < slot->used) slot = &g; }
        if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
        slot->is_loop = false;
        if (!c->no_cond_graph) {
            cudaGraph_t g = nullptr;
            cudaGraphConditionalHandle handle;
            bool ok = cudaGraphCreate(&g, 0) == cudaSuccess &&
                      cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault) == cudaSuccess;
            cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
            cudaGraphNode_t node;
            if (ok) {
                p.conditional.handle = handle; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
                ok = cudaGraphAddNode(&node, g, nullptr, 0, &p) == cudaSuccess;
            }
            if (ok) {
                cudaGraph_t body = p.conditional.phGraph_out[0];
                ok = cudaStreamBeginCaptureToGraph(c->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                if (ok) {
                    launch_correspond(mv, bv, cb, c->nn, grid_x, n_scans, c->stream, &me, &eb, grid_e);
                    for (int k = 0; k < lm; ++k) launch_evaluate(bv, cb, grid_x, n_scans, c->stream, &eb, grid_e);
                    launch_loop_cond(bv, n_scans, handle, c->stream);
                    cudaGraph_t dummy = nullptr;
                    ok = cudaStreamEndCapture(c->stream, &dummy) == cudaSuccess;
                }
            }
            if (ok) ok = cudaGraphInstantiate(&slot->exec, g, 0) == cudaSuccess;
            if (g) cudaGraphDestroy(g);
            if (ok) slot->is_loop = true;
            else { cudaGetLastError(); slot->exec = nullptr; c->no_cond_graph = true; }      // fall back to the unrolled schedule
        }
        if (!slot->exec) {
            cudaGraph_t g = nullptr;
            SO_CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            for (int it = 0; it < iters; ++it) {
                launch_correspond(mv, bv, cb, c->nn, grid_x, n_scans, c->stream, &me, &eb, grid_e);
                for (int k = 0; k < lm; ++k) launch_evaluate(bv, cb, grid_x, n_scans, c->stream, &eb, grid_e);
            }
            SO_CUDA_TRY(cudaStreamEndCapture(c->stream, &g));
            cudaError_t e = cudaGraphInstantiate(&slot->exec, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) { slot->exec = nullptr; return fail(SO_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
        }
        slot->key = key;
    }
    slot->used = ++c->graph_clock;
    SO_CUDA_TRY(cudaGraphLaunch(slot->exec, run_stream));
    *was_loop = slot->is_loop;
    if (!slot->is_loop) c->launches += uint64_t(iters) * (bv.counters && !grid_e ? kCorrLaunches + 1 + lm : kCorrLaunches + 2 + 2 * lm + (grid_e ? 1 + lm : 0));      // loop form: counted from the iterations executed
    return SO_OK;
}

static void init_state(IcpState& s, const double pose[7], uint32_t n, const so_icp_opts& o, uint32_t n_edge = 0) {
    std::memset(&s, 0, sizeof(s));
    std::memcpy(s.x0, pose, 7 * sizeof(double));
    std::memcpy(s.x, pose, 7 * sizeof(double));
    std::memcpy(s.cand, pose, 7 * sizeof(double));
    s.n_points = int32_t(n);
    s.n_input = int32_t(n);
    s.n_edge = int32_t(n_edge);
    s.max_icp_iters = o.max_icp_iters;
    s.lm_max_iterations = o.lm_max_iterations;
    // calculateSamplingRate (LidarSlam.cpp:346-351)
    s.sampling_rate = (o.max_surface_features > 0 && n > uint32_t(o.max_surface_features)) ? 1.0 * o.max_surface_features / double(n) : -1.0;
    s.use_prior = o.use_pose_prior ? 1 : 0;
    s.prior_vcf = o.visual_confidence_factor;
    for (int a = 0; a < 3; ++a) s.prior_unc[a] = o.prior_uncertainty[a];
    s.phase = PH_CORR;
}

static void fill_result(const Ctx* c, const IcpState& s, const double pose_in[7], const so_icp_opts& o, so_icp_result* r) {
    r->n_iterations = s.n_iterations;
    for (int i = 0; i < SO_MAX_ICP_ITERS; ++i) {
        r->iter_n_surf[i] = s.iter_n_surf[i]; r->iter_n_edge[i] = s.iter_n_edge[i]; r->iter_dtrans[i] = s.iter_dtrans[i]; r->iter_drot[i] = s.iter_drot[i];
        r->iter_lm_steps[i] = s.iter_lm_steps[i]; r->iter_lm_successful[i] = s.iter_lm_successful[i];
        r->iter_lm_termination[i] = s.iter_lm_termination[i]; r->iter_cost[i] = s.iter_cost[i];
    }
    std::memcpy(r->hist_obs, s.hist_obs, sizeof(r->hist_obs));
    std::memcpy(r->hist_reject_plane, s.hist_rej, sizeof(r->hist_reject_plane));
    std::memcpy(r->hist_reject_line, s.hist_rej_line, sizeof(r->hist_reject_line));
    std::memcpy(r->cov, s.cov, sizeof(r->cov));
    r->pos_err = s.pos_err; r->pos_inv_cond = s.pos_inv_cond; r->ori_err_deg = s.ori_err_deg; r->ori_inv_cond = s.ori_inv_cond;
    for (int i = 0; i < 3; ++i) { r->pos_dir[i] = s.pos_dir[i]; r->ori_dir[i] = s.ori_dir[i]; }
    if (s.status) r->status = s.status;
    r->knn_searched = s.knn_searched; r->knn_verified = s.knn_verified;
    r->prediction_source = s.use_prior ? 1 : 0;
    double T[7];
    std::memcpy(T, s.x, sizeof(T));
    std::memcpy(r->pose_opt, T, sizeof(T));
    // performPostOptimizationProcessing (LidarSlam.cpp:155-171): last_T_w_lidar == T_w_initial_guess == prior (:53-57)
    manual_yaw_correction(pose_in, T, double(o.yaw_ratio));
    std::memcpy(r->pose, T, sizeof(T));
    rel_motion(pose_in, T, &r->total_translation, &r->total_rotation);           // updateOptimizationStats (:198-210)
    r->translation_from_last = r->total_translation; r->rotation_from_last = r->total_rotation;
    (void)c;
}

// Shared tail of so_register / so_register_batch*: scans are on the device at d_scan (packed, back to back).
// d_scan: device scans (packed, back to back).  host_src != nullptr: the scans are still on the host (packed float4); this
// function uploads them into d_scan chunk by chunk on the copy stream.
static int register_core(Ctx* c, const float4* d_scan, const uint32_t* n_points, size_t n_scans, const double* poses,
                         const so_icp_opts* opts_in, so_icp_result* results, bool allow_shift, const void* host_src = nullptr,
                         const uint32_t* n_edge = nullptr, const uint32_t* d_inject = nullptr, int inject_iters = 0) {
    so_icp_opts o = *opts_in;
    if (o.lm_max_iterations <= 0) o.lm_max_iterations = 4;
    if (o.max_icp_iters <= 0 || o.max_icp_iters > SO_MAX_ICP_ITERS) return fail(SO_ERR_ARG, "max_icp_iters must be in [1,32]");
    if (o.lm_max_iterations > 16) return fail(SO_ERR_ARG, "lm_max_iterations must be <= 16");
    { int rc = ensure_maps(c); if (rc) return rc; }
    uint32_t max_n = 0, off = 0, eoff = 0;
    bool any = false, any_edge = false;
    for (size_t s = 0; s < n_scans; ++s) {
        so_icp_result* r = results + s;
        std::memset(r, 0, sizeof(*r));
        const double* pose = poses + 7 * s;
        std::memcpy(r->pose, pose, 7 * sizeof(double));
        std::memcpy(r->pose_opt, pose, 7 * sizeof(double));
        r->scan_surf_num = int32_t(n_points[s]);
        // prepareOptimizationState (LidarSlam.cpp:361-369)
        int32_t ijk[3];
        if (allow_shift && !o.skip_map_checks) { int rc = so_map_shift(reinterpret_cast<so_ctx*>(c), pose, ijk); if (rc < 0) return rc; }
        else for (int a = 0; a < 3; ++a) ijk[a] = block_coord_h(pose[a] + kHalfBlock, c->origin[a]);
        for (int a = 0; a < 3; ++a) r->pos_in_localmap[a] = ijk[a];
        r->map_surf_5x5 = counts_5x5(c, c->surf, ijk);
        r->map_edge_5x5 = counts_5x5(c, c->edge, ijk);
        init_state(c->h_state[s], pose, n_points[s], o, n_edge ? n_edge[s] : 0);
        c->h_offset[s] = off;
        off += n_points[s];
        c->h_eoffset[s] = eoff;                                    // edge clouds: packed back to back in d_escan like the scans in d_scan
        if (n_edge) { eoff += n_edge[s]; any_edge |= n_edge[s] != 0; r->scan_edge_num = int32_t(n_edge[s]); }
        if (!o.skip_map_checks && !(r->map_surf_5x5 > 50)) {       // hasEnoughFeatures (:379-381): pose stays the prior
            r->status = SO_STATUS_NOT_ENOUGH_FEATURES;
            c->h_state[s].phase = PH_DONE; c->h_state[s].status = r->status;
        } else if (n_points[s] == 0) {
            r->status = SO_STATUS_NO_CORRESPONDENCES;
            c->h_state[s].phase = PH_DONE; c->h_state[s].status = r->status;
        } else { any = true; max_n = std::max(max_n, n_points[s]); }
    }
    if (c->d_pose_sink && c->sink_cursor + n_scans > c->sink_cap) return fail(SO_ERR_CAPACITY, "pose sink full (so_set_pose_sink)");
    if (!any && c->d_pose_sink) {                               // nothing to register: the rows are the priors
        SO_CUDA_TRY(cudaMemcpyAsync(c->d_state, c->h_state, n_scans * sizeof(IcpState), cudaMemcpyHostToDevice, c->stream));
        launch_pack_poses(c->d_state, uint32_t(n_scans), c->d_pose_sink + c->sink_cursor * 8, c->stream);
        SO_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    if (any) {
        SO_CUDA_TRY(cudaMemcpyAsync(c->d_state, c->h_state, n_scans * sizeof(IcpState), cudaMemcpyHostToDevice, c->stream));
        SO_CUDA_TRY(cudaMemcpyAsync(c->d_offset, c->h_offset, n_scans * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        if (any_edge || c->eoffset_dirty) {                        // the all-zero table of the edge-less case is already there
            SO_CUDA_TRY(cudaMemcpyAsync(c->d_eoffset, c->h_eoffset, n_scans * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
            c->eoffset_dirty = any_edge;
        }
        count_h2d(c, n_scans * (sizeof(IcpState) + sizeof(uint32_t)));
        // Chunk the batch.  Even chunks run on `stream`, odd ones on aux_stream, so that one chunk's serial stretches (the
        // one-CTA optimiser steps, kernel tails, the loop condition) sit under the other's wide kernels; with host input
        // the H2D of chunk k+1 (copy stream) also overlaps the kernels of chunk k.
        // Device-resident scans: two chunks (one per stream).  Chunks stay >= 8 scans: below that the wide kernels stop filling the GPU.
        // (the rule -- growing chunks for host input -- lives in so_chunks.h, where a CPU test can reach it)
        const std::vector<uint32_t> bounds = chunk_bounds(n_scans, host_src != nullptr, c->profiling, c->chunk_override);
        const size_t n_chunks = bounds.size() - 1;
        std::vector<Chunk> chunks;
        for (size_t k = 0; k < n_chunks; ++k) {
            const uint32_t f = bounds[k], e = bounds[k + 1];
            if (e == f) continue;
            Chunk ch{f, e - f, c->h_offset[f], (e < n_scans ? c->h_offset[e] : off) - c->h_offset[f], 0};
            uint32_t mx = 0, mx_in = 0;
            for (uint32_t s = f; s < e; ++s) if (c->h_state[s].phase == PH_CORR) {
                mx_in = std::max(mx_in, n_points[s]);
                // shouldProcessPoint keeps at most max_surface_features + 1 points of a decimated scan (one per wrap of i * rate)
                const uint32_t kept = c->h_state[s].sampling_rate < 0.0 ? n_points[s] : std::min<uint32_t>(n_points[s], uint32_t(o.max_surface_features) + 1);
                mx = std::max(mx, kept);
            }
            ch.grid_in = (mx_in + kThreads - 1) / kThreads;
            ch.small = mx <= kPrepareSmallCap && mx_in > 0;
            ch.max_kept = mx;
            ch.grid_x = (mx + kThreads - 1) / kThreads;
            // rounded up to a bucket of 16 CTAs so that scans of slightly different sizes (live SLAM: every scan differs) share one
            // captured graph; the kernels guard i < n_points and k_lm_step sums only the partial rows n_points implies
            if (ch.grid_x) ch.grid_x = std::min<uint32_t>(c->grid_x_cap, (ch.grid_x + 15u) & ~15u);
            if (n_edge) for (uint32_t s = f; s < e; ++s) if (c->h_state[s].phase == PH_CORR) ch.grid_e = std::max(ch.grid_e, (n_edge[s] + kThreads - 1) / kThreads);
            chunks.push_back(ch);
        }
        SO_CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
        const bool two_streams = chunks.size() > 1 && !c->single_stream;
        if (two_streams) {
            SO_CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));                    // aux starts after the state upload and all earlier work
            SO_CUDA_TRY(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
        }
        if (host_src) {
            SO_CUDA_TRY(cudaEventRecord(c->ev_copy[16], c->stream));                // copies must not overtake earlier work on d_scan
            SO_CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_copy[16], 0));
            for (size_t k = 0; k < chunks.size(); ++k) {
                const Chunk& ch = chunks[k];
                if (ch.pt_count)
                    SO_CUDA_TRY(cudaMemcpyAsync(c->d_scan + ch.pt_first, static_cast<const float4*>(host_src) + ch.pt_first,
                                                size_t(ch.pt_count) * sizeof(float4), cudaMemcpyHostToDevice, c->copy_stream));
                count_h2d(c, size_t(ch.pt_count) * sizeof(float4));
                SO_CUDA_TRY(cudaEventRecord(c->ev_copy[k], c->copy_stream));
            }
        }
        std::vector<char> loop_flags(chunks.size(), 0);
        for (size_t k = 0; k < chunks.size(); ++k) {
            const Chunk& ch = chunks[k];
            cudaStream_t st = (two_streams && (k & 1)) ? c->aux_stream : c->stream;
            if (host_src) SO_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_copy[k], 0));
            if (ch.grid_x == 0) continue;
            int rc = prepare_scans(c, d_scan, ch, st, true);
            bool was_loop = false;
            if (!rc) rc = run_schedule(c, ch, o.max_icp_iters, o.lm_max_iterations, false, &was_loop, st, d_inject, inject_iters);
            if (rc) {                                       // leave no work behind on the side streams before reporting the error
                cudaStreamSynchronize(c->aux_stream);
                cudaStreamSynchronize(c->copy_stream);
                return rc;
            }
            loop_flags[k] = was_loop;
        }
        if (two_streams) {
            SO_CUDA_TRY(cudaEventRecord(c->ev_join, c->aux_stream));
            SO_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        }
        SO_CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
        { int rc = run_deferred_upload(c); if (rc) return rc; }      // host-side staging of the whole cloud while the registration runs
        if (c->d_pose_sink) { launch_pack_poses(c->d_state, uint32_t(n_scans), c->d_pose_sink + c->sink_cursor * 8, c->stream); c->launches++; }
        SO_CUDA_TRY(cudaMemcpyAsync(c->h_state, c->d_state, n_scans * sizeof(IcpState), cudaMemcpyDeviceToHost, c->stream));
        count_d2h(c, n_scans * sizeof(IcpState));
        SO_CUDA_TRY(cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        for (size_t k = 0; k < chunks.size(); ++k) {
            if (!loop_flags[k]) continue;
            int max_it = 0;
            for (uint32_t s = chunks[k].first; s < chunks[k].first + chunks[k].count; ++s) max_it = std::max(max_it, int(c->h_state[s].n_iterations));
            const bool fused = chunks[k].count <= 2 && !chunks[k].grid_e && !c->no_fused_lm;
            c->launches += uint64_t(max_it) * uint64_t(fused ? kCorrLaunches + 2 + o.lm_max_iterations
                                                             : kCorrLaunches + 3 + 2 * o.lm_max_iterations + (chunks[k].grid_e ? 1 + o.lm_max_iterations : 0));
        }
        for (size_t s = 0; s < n_scans; ++s) {
            if (results[s].status == SO_STATUS_NOT_ENOUGH_FEATURES || n_points[s] == 0) continue;
            fill_result(c, c->h_state[s], poses + 7 * s, o, results + s);
            results[s].time_ms = double(ms);
        }
    }
    if (c->d_pose_sink) c->sink_cursor += n_scans;
    return SO_OK;
}

}  // namespace so

// =================================================================================================== C ABI
using namespace so;

extern "C" {

const char* so_last_error(void) { return g_err.c_str(); }

int so_device_available(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return 0; }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { cudaGetLastError(); return 0; }
    return p.major == 10 ? 1 : 0;
}

so_ctx* so_create(const so_config* cfg_in) {
    so_config cfg{};
    if (cfg_in) cfg = *cfg_in;
    if (cfg.max_map_points == 0) cfg.max_map_points = 4u << 20;
    if (cfg.max_scan_points == 0) cfg.max_scan_points = 262144;
    if (cfg.max_batch == 0) cfg.max_batch = 1;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); set_error("no CUDA device: this library has no CPU fallback"); return nullptr; }
    if (cfg.device < 0 || cfg.device >= n) { set_error("bad device ordinal"); return nullptr; }
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, cfg.device);
    if (p.major != 10) { set_error("device is not compute capability 10.x (kernels are built for sm_100a only)"); return nullptr; }
    Ctx* c = new Ctx();
    c->device = cfg.device;
    c->cfg = cfg;
    if (cfg.plane_res > 0) c->surf.res = cfg.plane_res;
    if (cfg.line_res > 0) c->edge.res = cfg.line_res;
    if (std::getenv("SO_NO_COND_GRAPH")) c->no_cond_graph = true;
    if (std::getenv("SO_SINGLE_STREAM")) c->single_stream = true;
    if (std::getenv("SO_NO_FUSED_LM")) c->no_fused_lm = true;
    if (std::getenv("SO_NO_SMALL_PREPARE")) c->no_small_prepare = true;
    if (std::getenv("SO_NO_COOP_KNN")) c->no_coop_knn = true;
    if (std::getenv("SO_FORCE_KEY64")) c->force_key64 = true;
    if (std::getenv("SO_NO_DEC_UPLOAD")) c->no_dec_upload = true;
    if (const char* e = std::getenv("SO_CHUNKS")) c->chunk_override = std::max(0, std::min(16, std::atoi(e)));      // profiling aid: ncu cannot see inside conditional-node bodies
    if (ctx_alloc(c) != SO_OK) { ctx_free(c); return nullptr; }
    return reinterpret_cast<so_ctx*>(c);
}

void so_destroy(so_ctx* ctx) { ctx_free(reinterpret_cast<Ctx*>(ctx)); }

int so_set_stream(so_ctx* ctx, void* cuda_stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return fail(SO_ERR_ARG, "null ctx");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    SO_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (auto& g : c->graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    if (cuda_stream) {
        if (c->own_stream) cudaStreamDestroy(c->stream);
        c->stream = static_cast<cudaStream_t>(cuda_stream); c->own_stream = false;
    } else if (!c->own_stream) {
        SO_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true;
    }
    return SO_OK;
}

// ---- map -------------------------------------------------------------------------------------------------------
int so_map_set_resolution(so_ctx* ctx, float line_res, float plane_res) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !(plane_res > 0)) return fail(SO_ERR_ARG, "bad args");
    if (line_res > 0 && line_res != c->edge.res) {
        const int nb_old = c->edge.nb;
        c->edge.res = line_res;
        c->map_epoch++;
        if (map_cells_per_block(line_res) != nb_old) c->edge.dirty = true;
    }
    if (plane_res != c->surf.res) {
        const int nb_old = c->surf.nb;
        c->surf.res = plane_res;
        c->map_epoch++;                                     // bound_d2 / plane_res live in the captured MapView
        if (map_cells_per_block(plane_res) != nb_old) c->surf.dirty = true;
    }
    return SO_OK;
}

int so_map_set_origin(so_ctx* ctx, const double t[3], int32_t out_origin[3]) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !t) return fail(SO_ERR_ARG, "bad args");
    for (int a = 0; a < 3; ++a) {                           // LocalMap::setOrigin (LocalMap.h:146-164)
        int cc = int((t[a] + kHalfBlock) / kBlock);
        if (t[a] + kHalfBlock < 0) cc--;
        c->origin[a] = -cc;
        if (out_origin) out_origin[a] = c->origin[a];
    }
    c->surf.dirty = true; c->edge.dirty = true;
    return SO_OK;
}

int so_map_get_origin(so_ctx* ctx, int32_t out[3]) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !out) return fail(SO_ERR_ARG, "bad args");
    for (int a = 0; a < 3; ++a) out[a] = c->origin[a];
    return SO_OK;
}

int so_map_shift(so_ctx* ctx, const double t[3], int32_t out_ijk[3]) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !t || !out_ijk) return fail(SO_ERR_ARG, "bad args");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    int cc[3], shift[3] = {0, 0, 0};
    const int dims[3] = {kW, kH, kD};
    for (int a = 0; a < 3; ++a) {                           // LocalMap::shiftMap (LocalMap.h:169-287)
        cc[a] = block_coord_h(t[a] + kHalfBlock, c->origin[a]);
        while (cc[a] < 3) { cc[a]++; shift[a]++; }
        while (cc[a] >= dims[a] - 3) { cc[a]--; shift[a]--; }
    }
    if (shift[0] || shift[1] || shift[2]) {
        for (int a = 0; a < 3; ++a) c->origin[a] += shift[a];
        c->surf.dirty = true; c->edge.dirty = true;
    }
    { int rc = ensure_maps(c); if (rc) return rc; }
    for (int a = 0; a < 3; ++a) out_ijk[a] = cc[a];
    return SO_OK;
}

static int store_set_points(Ctx* c, MapStore& ms, const void* xyzi, size_t n, size_t stride, size_t ioff) {
    if (!c || (!xyzi && n) || stride < 12) return fail(SO_ERR_ARG, "bad args");
    if (n > c->cfg.max_map_points) return fail(SO_ERR_CAPACITY, "map larger than so_config.max_map_points");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    int rc = upload_cloud(c, xyzi, n, stride, ioff, ms.d_xyzi);
    if (rc) return rc;
    ms.n = uint32_t(n);
    return map_rebuild(c, ms);
}
// xyzi == nullptr with n != 0: the n points of the scan last uploaded by so_register (c->d_scan)
static int store_add(Ctx* c, MapStore& ms, const void* xyzi, size_t n, size_t stride, size_t ioff, const double* pose) {
    if (!c || stride < 12) return fail(SO_ERR_ARG, "bad args");
    if (size_t(ms.n) + n > c->cfg.max_map_points) return fail(SO_ERR_CAPACITY, "map + new points exceed so_config.max_map_points");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    if (n == 0) return SO_OK;
    int rc = SO_OK;
    if (ms.dirty) { rc = map_rebuild(c, ms); if (rc) return rc; }           // pending origin / resolution change first: it may compact the cloud
    if (xyzi) rc = upload_cloud(c, xyzi, n, stride, ioff, ms.d_xyzi + ms.n);
    else SO_CUDA_TRY(cudaMemcpyAsync(ms.d_xyzi + ms.n, c->d_scan, n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));   // the registered scan, still on the device
    if (rc) return rc;
    if (pose) { rc = map_transform_tail(c, ms, uint32_t(n), pose); if (rc) return rc; }
    timed_launch_begin(c);
    rc = map_add_points(c, ms, uint32_t(n));
    timed_launch_end(c, 3);
    return rc;
}

int so_map_set_points(so_ctx* ctx, const void* xyzi, size_t n, size_t stride, size_t ioff) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    return c ? store_set_points(c, c->surf, xyzi, n, stride, ioff) : fail(SO_ERR_ARG, "null ctx");
}
int so_map_set_edge_points(so_ctx* ctx, const void* xyzi, size_t n, size_t stride, size_t ioff) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    return c ? store_set_points(c, c->edge, xyzi, n, stride, ioff) : fail(SO_ERR_ARG, "null ctx");
}
int so_map_add_surf(so_ctx* ctx, const void* xyzi, size_t n, size_t stride, size_t ioff) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!xyzi && n) return fail(SO_ERR_ARG, "bad args");
    return c ? store_add(c, c->surf, xyzi, n, stride, ioff, nullptr) : fail(SO_ERR_ARG, "null ctx");
}
int so_map_add_edge(so_ctx* ctx, const void* xyzi, size_t n, size_t stride, size_t ioff) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!xyzi && n) return fail(SO_ERR_ARG, "bad args");
    return c ? store_add(c, c->edge, xyzi, n, stride, ioff, nullptr) : fail(SO_ERR_ARG, "null ctx");
}
int so_map_add_scan(so_ctx* ctx, const void* xyzi, size_t n, size_t stride, size_t ioff, const double pose[7]) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!pose || (!xyzi && n)) return fail(SO_ERR_ARG, "bad args");
    return c ? store_add(c, c->surf, xyzi, n, stride, ioff, pose) : fail(SO_ERR_ARG, "null ctx");
}
int so_map_add_registered_scan(so_ctx* ctx, const double pose[7]) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !pose) return fail(SO_ERR_ARG, "bad args");
    if (c->last_scan_n == 0) return SO_OK;
    return store_add(c, c->surf, nullptr, c->last_scan_n, 16, 12, pose);
}
int so_map_add_scan_edge(so_ctx* ctx, const void* xyzi, size_t n, size_t stride, size_t ioff, const double pose[7]) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!pose || (!xyzi && n)) return fail(SO_ERR_ARG, "bad args");
    return c ? store_add(c, c->edge, xyzi, n, stride, ioff, pose) : fail(SO_ERR_ARG, "null ctx");
}

int so_map_counts_5x5(so_ctx* ctx, const int32_t ijk[3], int32_t* n_edge, int32_t* n_surf) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !ijk) return fail(SO_ERR_ARG, "bad args");
    { int rc = ensure_maps(c); if (rc) return rc; }
    if (n_edge) *n_edge = counts_5x5(c, c->edge, ijk);
    if (n_surf) *n_surf = counts_5x5(c, c->surf, ijk);
    return SO_OK;
}

size_t so_map_size(so_ctx* ctx) { Ctx* c = reinterpret_cast<Ctx*>(ctx); return c ? size_t(c->surf.n) + size_t(c->edge.n) : 0; }

int so_map_download(so_ctx* ctx, int mode, const int32_t ijk[3], float* out, size_t cap, size_t* n_out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !n_out) return fail(SO_ERR_ARG, "bad args");
    if (mode == 1 && !ijk) return fail(SO_ERR_ARG, "mode 1 needs ijk");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    { int rc = ensure_maps(c); if (rc) return rc; }
    // LocalMap::getAllLocalMap / get5x5LocalMap (LocalMap.h:647-687): cubes in index order, each cube's edge cloud then its surf cloud
    std::vector<float4> he(c->edge.n), hs(c->surf.n);
    if (c->edge.n) SO_CUDA_TRY(cudaMemcpy(he.data(), c->edge.d_xyzi, size_t(c->edge.n) * sizeof(float4), cudaMemcpyDeviceToHost));
    if (c->surf.n) SO_CUDA_TRY(cudaMemcpy(hs.data(), c->surf.d_xyzi, size_t(c->surf.n) * sizeof(float4), cudaMemcpyDeviceToHost));
    struct Ref { int32_t lin; uint8_t kind; uint32_t idx; };
    std::vector<Ref> refs;
    refs.reserve(he.size() + hs.size());
    auto push = [&](const std::vector<float4>& h, uint8_t kind) {
        for (uint32_t i = 0; i < h.size(); ++i) {
            int g[3]; const float q[3] = {h[i].x, h[i].y, h[i].z};
            for (int a = 0; a < 3; ++a) g[a] = block_coord_h(double(q[a]) + kHalfBlock, c->origin[a]);
            if (mode == 1 && (std::abs(g[0] - ijk[0]) > 2 || std::abs(g[1] - ijk[1]) > 2 || std::abs(g[2] - ijk[2]) > 1)) continue;
            refs.push_back(Ref{g[0] + kW * g[1] + kW * kH * g[2], kind, i});
        }
    };
    push(he, 0); push(hs, 1);
    if (mode == 1) {    // get5x5LocalMap loops i (x) outermost, then j, then k: order by (i, j, k), not by the linear cube index
        std::stable_sort(refs.begin(), refs.end(), [](const Ref& a, const Ref& b) {
            const int ai = a.lin % kW, aj = (a.lin / kW) % kH, ak = a.lin / (kW * kH), bi = b.lin % kW, bj = (b.lin / kW) % kH, bk = b.lin / (kW * kH);
            if (ai != bi) return ai < bi; if (aj != bj) return aj < bj; if (ak != bk) return ak < bk; return a.kind < b.kind; });
    } else {
        std::stable_sort(refs.begin(), refs.end(), [](const Ref& a, const Ref& b) { return a.lin != b.lin ? a.lin < b.lin : a.kind < b.kind; });
    }
    size_t k = 0;
    for (const Ref& r : refs) {
        if (out && k < cap) std::memcpy(out + 4 * k, r.kind ? &hs[r.idx] : &he[r.idx], sizeof(float4));
        ++k;
    }
    *n_out = k;
    return SO_OK;
}

// ---- scan pre-filter ---------------------------------------------------------------------------------------------
int so_scan_prefilter(so_ctx* ctx, const void* xyzi, size_t n, size_t stride, size_t ioff, int auto_voxel_size, float* line_res,
                      float* plane_res, float* out_xyzi, size_t cap, size_t* n_out, double* average_distance) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || (!xyzi && n) || stride < 12 || !line_res || !plane_res || !n_out || (!out_xyzi && cap)) return fail(SO_ERR_ARG, "bad args");
    if (n > c->scan_cap) return fail(SO_ERR_CAPACITY, "cloud larger than the scan buffers");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    *n_out = 0;
    if (auto_voxel_size && n) {
        // laserMapping::adjustVoxelSize (laserMapping.cpp:603-636): float accumulators in cloud order, exactly as written there
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        int far_points = 0;
        const unsigned char* p = static_cast<const unsigned char*>(xyzi);
        for (size_t i = 0; i < n; ++i, p += stride) {
            float v[3];
            std::memcpy(v, p, 12);
            a0 += std::fabs(v[0]); a1 += std::fabs(v[1]); a2 += std::fabs(v[2]);
            if (v[0] * v[0] + v[1] * v[1] + v[2] * v[2] > 9) far_points++;
        }
        (void)far_points;                                          // increase_blind_radius is computed and never used upstream
        const float sz = float(n);
        a0 /= sz; a1 /= sz; a2 /= sz;
        const double avg = double(a0 * a1 * a2);                   // float product stored into a float64 field
        if (average_distance) *average_distance = avg;
        if (avg < 25) { *line_res = 0.1f; *plane_res = 0.2f; }
        else if (avg > 65) { *line_res = 0.4f; *plane_res = 0.8f; }
    }
    int rc = so_map_set_resolution(ctx, *line_res, *plane_res);   // slam.localMap.lineRes_/planeRes_ = config_ (:648-649)
    if (rc) return rc;
    rc = upload_cloud(c, xyzi, n, stride, ioff, c->d_scan);
    if (rc) return rc;
    uint32_t m = 0;
    timed_launch_begin(c);
    rc = scan_voxel_filter(c, uint32_t(n), *plane_res, &m);
    timed_launch_end(c, 3);
    if (rc) return rc;
    *n_out = m;
    c->prefiltered_n = m;                                          // so_register_prefiltered: the filtered cloud stays in d_scan_sorted
    const size_t ncopy = std::min<size_t>(m, cap);
    if (ncopy) {
        SO_CUDA_TRY(cudaMemcpyAsync(out_xyzi, c->d_scan_sorted, ncopy * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
        count_d2h(c, ncopy * sizeof(float4));
        SO_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return SO_OK;
}

// ---- scan preparation: deskew + uniform extraction (featureExtraction.cpp:222-314,504-525) ----------------------
namespace {
struct HPose { double q[4]; double p[3]; };      // xyzw
HPose hpose_from7(const double* v) { return HPose{{v[3], v[4], v[5], v[6]}, {v[0], v[1], v[2]}}; }
void hnormalize(double q[4]) {
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; ++k) q[k] /= n;
}
// Eigen::QuaternionBase::slerp
void hslerp(const double a[4], double t, const double b[4], double o[4]) {
    const double one = 1.0 - std::numeric_limits<double>::epsilon();
    const double d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
    const double ad = std::fabs(d);
    double s0, s1;
    if (ad >= one) { s0 = 1.0 - t; s1 = t; }
    else {
        const double th = std::acos(ad), st = std::sin(th);
        s0 = std::sin((1.0 - t) * th) / st;
        s1 = std::sin(t * th) / st;
    }
    if (d < 0) s1 = -s1;
    for (int k = 0; k < 4; ++k) o[k] = s0 * a[k] + s1 * b[k];
}
// the sign Eigen's matrix -> quaternion conversion produces (w > 0 when the trace is positive, else the dominant axis > 0)
void hcanonical_sign(double q[4]) {
    const double tr = 4.0 * q[3] * q[3] - 1.0;        // trace of R(q) for a unit q
    bool flip;
    if (tr > 0) flip = q[3] < 0;
    else {
        int i = 0;
        if (std::fabs(q[1]) > std::fabs(q[0])) i = 1;
        if (std::fabs(q[2]) > std::fabs(q[i])) i = 2;
        flip = q[i] < 0;
    }
    if (flip) for (int k = 0; k < 4; ++k) q[k] = -q[k];
}
}  // namespace

int so_scan_deskew(so_ctx* ctx, void* points, size_t n, size_t stride, size_t time_offset, double lidar_start_time, const double* sample_times,
                   const double* sample_poses, size_t n_samples, int imu_only, const double T_i_l[7], double start_pose_out[7], size_t* n_past_end) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || (!points && n) || stride < 12 || time_offset + 4 > stride || !sample_times || !sample_poses || n_samples == 0 || (imu_only && !T_i_l))
        return fail(SO_ERR_ARG, "bad args");
    if (n > c->scan_cap) return fail(SO_ERR_CAPACITY, "cloud larger than the scan buffers");
    if (n_samples * 8 > c->scan_cap) return fail(SO_ERR_CAPACITY, "more pose samples than the scratch holds (max_scan_points / 8)");
    for (size_t k = 1; k < n_samples; ++k)
        if (!(sample_times[k] > sample_times[k - 1])) return fail(SO_ERR_ARG, "sample_times must be strictly ascending (std::map keys)");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    // getInterpolatedPoseAtTime(lidar_start_time) on the host (:279): same rule as the kernel
    size_t after = size_t(std::upper_bound(sample_times, sample_times + n_samples, lidar_start_time) - sample_times);
    if (after == n_samples) after = n_samples - 1;
    if (sample_times[after] < 0.0001) after = 0;
    HPose start;
    if (after == 0) start = hpose_from7(sample_poses);
    else {
        const HPose a = hpose_from7(sample_poses + 7 * (after - 1)), b = hpose_from7(sample_poses + 7 * after);
        const double ratio = (lidar_start_time - sample_times[after - 1]) / (sample_times[after] - sample_times[after - 1]);
        hslerp(a.q, ratio, b.q, start.q);
        for (int k = 0; k < 3; ++k) start.p[k] = (1 - ratio) * a.p[k] + ratio * b.p[k];
    }
    if (imu_only) start.p[0] = start.p[1] = start.p[2] = 0.0;          // Imu::Ptr samples carry no translation (:232-236)
    hnormalize(start.q);
    DeskewParams P{};
    P.start_time = lidar_start_time;
    P.n_samples = uint32_t(n_samples);
    P.n_smem = uint32_t(std::min<size_t>(n_samples, 2048));
    P.imu_only = imu_only ? 1 : 0;
    P.q0_conj[0] = -start.q[0]; P.q0_conj[1] = -start.q[1]; P.q0_conj[2] = -start.q[2]; P.q0_conj[3] = start.q[3];
    for (int k = 0; k < 3; ++k) P.p0[k] = start.p[k];
    double out7[7] = {start.p[0], start.p[1], start.p[2], start.q[0], start.q[1], start.q[2], start.q[3]};
    if (imu_only) {
        HPose il = hpose_from7(T_i_l);
        hnormalize(il.q);
        for (int k = 0; k < 4; ++k) P.q_il[k] = il.q[k];
        for (int k = 0; k < 3; ++k) P.t_il[k] = il.p[k];
        P.q_li[0] = -il.q[0]; P.q_li[1] = -il.q[1]; P.q_li[2] = -il.q[2]; P.q_li[3] = il.q[3];
        double r[3];
        qrot(P.q_li, il.p, r);                                           // Twist::inverse (Twist.h:172-179)
        P.t_li[0] = -r[0]; P.t_li[1] = -r[1]; P.t_li[2] = -r[2];
        // T_w_original_sensor = T_w_original * T_i_l (:283-286)
        double q[4];
        qmul(start.q, il.q, q);
        hnormalize(q);
        hcanonical_sign(q);
        qrot(start.q, il.p, r);
        out7[0] = r[0] + start.p[0]; out7[1] = r[1] + start.p[1]; out7[2] = r[2] + start.p[2];
        out7[3] = q[0]; out7[4] = q[1]; out7[5] = q[2]; out7[6] = q[3];
    }
    if (start_pose_out) std::memcpy(start_pose_out, out7, sizeof(out7));
    if (n_past_end) *n_past_end = 0;
    if (n == 0) return SO_OK;
    int rc = upload_cloud(c, points, n, stride, time_offset, c->d_scan);          // float4 {x, y, z, time}
    if (rc) return rc;
    // pose samples into device scratch (the scan-sort key buffer): [times n][poses 7n]; IMU samples lose their translation here
    std::vector<double> hs(n_samples * 8);
    for (size_t k = 0; k < n_samples; ++k) {
        hs[k] = sample_times[k];
        for (int j = 0; j < 7; ++j) hs[n_samples + 7 * k + j] = (imu_only && j < 3) ? 0.0 : sample_poses[7 * k + j];
    }
    double* d_samples = reinterpret_cast<double*>(c->d_skeys);
    SO_CUDA_TRY(cudaMemcpyAsync(d_samples, hs.data(), hs.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    count_h2d(c, hs.size() * sizeof(double));
    SO_CUDA_TRY(cudaMemsetAsync(c->d_svals, 0, 4, c->stream));
    P.times = d_samples; P.poses = d_samples + n_samples; P.past_end = c->d_svals;
    timed_launch_begin(c);
    launch_deskew(c->d_scan, uint32_t(n), P, c->stream);
    c->launches++;
    timed_launch_end(c, 3);
    rc = ensure_stage(c, n * sizeof(float4));
    if (rc) return rc;
    uint32_t past = 0;
    SO_CUDA_TRY(cudaMemcpyAsync(c->h_stage, c->d_scan, n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(&past, c->d_svals, 4, cudaMemcpyDeviceToHost, c->stream));
    count_d2h(c, n * sizeof(float4) + 4);
    SO_CUDA_TRY(cudaStreamSynchronize(c->stream));                                 // also covers hs going out of scope
    const float4* h = static_cast<const float4*>(c->h_stage);
    unsigned char* dst = static_cast<unsigned char*>(points);
    for (size_t i = 0; i < n; ++i, dst += stride) std::memcpy(dst, &h[i], 12);     // x, y, z only; every other field untouched
    if (n_past_end) *n_past_end = past;
    return SO_OK;
}

int so_scan_extract_uniform(so_ctx* ctx, const void* points, size_t n, size_t stride, size_t time_offset, int skip_num, float block_range,
                            int int_abs, float* out_xyzi, size_t cap, size_t* n_out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || (!points && n) || stride < 12 || time_offset + 4 > stride || !n_out || (!out_xyzi && cap)) return fail(SO_ERR_ARG, "bad args");
    if (skip_num <= 0) return fail(SO_ERR_ARG, "skip_num must be positive (the reference loop would not terminate)");
    if (n > c->scan_cap) return fail(SO_ERR_CAPACITY, "cloud larger than the scan buffers");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    *n_out = 0;
    int rc = upload_cloud(c, points, n, stride, time_offset, c->d_scan);
    if (rc) return rc;
    uint32_t m = 0;
    timed_launch_begin(c);
    rc = scan_extract_uniform(c, uint32_t(n), uint32_t(skip_num), block_range, int_abs, &m);
    timed_launch_end(c, 3);
    if (rc) return rc;
    *n_out = m;
    const size_t ncopy = std::min<size_t>(m, cap);
    if (ncopy) {
        SO_CUDA_TRY(cudaMemcpyAsync(out_xyzi, c->d_scan_sorted, ncopy * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
        count_d2h(c, ncopy * sizeof(float4));
        SO_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return SO_OK;
}

// ---- registration ----------------------------------------------------------------------------------------------
int so_register(so_ctx* ctx, const void* surf, size_t n_surf, const void* edge, size_t n_edge, size_t stride, size_t ioff,
                const double pose_in[7], const so_icp_opts* opts, so_icp_result* out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !pose_in || !opts || !out || (!surf && n_surf) || (!edge && n_edge) || stride < 12) return fail(SO_ERR_ARG, "bad args");
    if (n_edge > c->edge_cap) return fail(SO_ERR_CAPACITY, "edge cloud larger than so_config.max_scan_points");
    if (n_surf > c->cfg.max_scan_points) return fail(SO_ERR_CAPACITY, "scan larger than so_config.max_scan_points");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    const uint32_t n = uint32_t(n_surf);
    const uint32_t ne = uint32_t(n_edge);
    int rc;
    // Decimated scan (the shipped configurations cap a 28 800-point scan at 2 000 features): the registration reads only the points
    // shouldProcessPoint keeps, so only those go up in front of it (32 KB instead of 460 KB); the whole cloud -- needed by
    // so_map_add_registered_scan, not by the registration -- is staged and copied on the copy stream while the registration runs.
    const bool decimated = !c->no_dec_upload && opts->max_surface_features > 0 && n > uint32_t(opts->max_surface_features) &&
                           uint32_t(opts->max_surface_features) + 1 <= kPrepareSmallCap;
    if (decimated) {
        const std::vector<uint32_t>& keep = decimation_list(c, n, opts->max_surface_features);
        const uint32_t kept = uint32_t(keep.size());
        if (kept == 0 || kept > kPrepareSmallCap) return fail(SO_ERR_ARG, "decimation list out of range");
        const bool packed = stride == 16 && ioff == 12;
        rc = upload_cloud(c, edge, n_edge, stride, ioff, c->d_escan);   // first: a strided edge cloud goes through the same staging buffer
        if (rc) return rc;
        if (n_edge && !packed) SO_CUDA_TRY(cudaStreamSynchronize(c->stream));       // ... and has to have left it
        SO_CUDA_TRY(cudaEventSynchronize(c->ev_dec));                    // so has the previous call's deferred copy
        rc = ensure_stage(c, (size_t(kPrepareSmallCap) + (packed ? 0 : n_surf)) * sizeof(float4));
        if (rc) return rc;
        float4* h = static_cast<float4*>(c->h_stage);
        const unsigned char* p = static_cast<const unsigned char*>(surf);
        for (uint32_t j = 0; j < kept; ++j) h[j] = read_point(p + size_t(keep[j]) * stride, stride, ioff);
        SO_CUDA_TRY(cudaEventRecord(c->ev_copy[16], c->stream));        // the whole-cloud copy must not overtake earlier readers of d_scan
        SO_CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_copy[16], 0));
        SO_CUDA_TRY(cudaMemcpyAsync(c->d_dec, h, size_t(kept) * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
        c->bytes_h2d += size_t(kept) * sizeof(float4);
        c->defer.src = surf; c->defer.n = n_surf; c->defer.stride = stride; c->defer.ioff = ioff; c->defer.stage_off = kPrepareSmallCap;
        c->last_scan_n = n;
        so_icp_opts o2 = *opts;
        o2.max_surface_features = 0;                                     // the decimation has been applied
        rc = register_core(c, c->d_dec, &kept, 1, pose_in, &o2, out, true, nullptr, &ne);
        const int rc2 = run_deferred_upload(c);                          // not yet run when nothing was launched (soft statuses) or on errors
        SO_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_dec, 0));       // later work on the context stream (the insert) sees the whole cloud
        if (rc) return rc;
        if (rc2) return rc2;
        out->scan_surf_num = int32_t(n_surf);
    } else {
        rc = upload_cloud(c, surf, n_surf, stride, ioff, c->d_scan);
        if (rc) return rc;
        c->last_scan_n = n;                                             // so_map_add_registered_scan inserts it without a second upload
        rc = upload_cloud(c, edge, n_edge, stride, ioff, c->d_escan);   // edge branch input (empty upstream: featureExtraction.cpp:429-436)
        if (rc) return rc;
        rc = register_core(c, c->d_scan, &n, 1, pose_in, opts, out, true, nullptr, &ne);
        if (rc) return rc;
    }
    out->scan_edge_num = int32_t(n_edge);
    out->time_total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return out->status;
}

int so_register_prefiltered(so_ctx* ctx, const double pose_in[7], const so_icp_opts* opts, so_icp_result* out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !pose_in || !opts || !out) return fail(SO_ERR_ARG, "bad args");
    if (c->prefiltered_n == SIZE_MAX) return fail(SO_ERR_ARG, "so_register_prefiltered needs a preceding so_scan_prefilter (and nothing that reuses the scan buffers in between)");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    const uint32_t n = uint32_t(c->prefiltered_n);
    c->prefiltered_n = SIZE_MAX;
    if (n) SO_CUDA_TRY(cudaMemcpyAsync(c->d_scan, c->d_scan_sorted, size_t(n) * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    c->last_scan_n = n;
    int rc = register_core(c, c->d_scan, &n, 1, pose_in, opts, out, true);
    if (rc) return rc;
    out->scan_edge_num = 0;
    out->time_total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return out->status;
}

int so_register_injected(so_ctx* ctx, const void* surf, size_t n_surf, size_t stride, size_t ioff, const double pose_in[7], const so_icp_opts* opts,
                         const uint32_t* nn_ids, int32_t n_trace_iters, so_icp_result* out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !pose_in || !opts || !out || !surf || !nn_ids || n_surf == 0 || n_trace_iters <= 0 || stride < 12) return fail(SO_ERR_ARG, "bad args");
    if (n_surf > c->cfg.max_scan_points) return fail(SO_ERR_CAPACITY, "scan larger than so_config.max_scan_points");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    const size_t words = size_t(n_trace_iters) * n_surf * 5;
    if (words > c->inject_cap) {
        cudaFree(c->d_inject); c->d_inject = nullptr; c->inject_cap = 0;
        SO_CUDA_TRY(cudaMalloc(&c->d_inject, words * sizeof(uint32_t)));
        c->inject_cap = words;
    }
    SO_CUDA_TRY(cudaMemcpyAsync(c->d_inject, nn_ids, words * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    int rc = upload_cloud(c, surf, n_surf, stride, ioff, c->d_scan);
    if (rc) return rc;
    const uint32_t n = uint32_t(n_surf);
    rc = register_core(c, c->d_scan, &n, 1, pose_in, opts, out, true, nullptr, nullptr, c->d_inject, int(n_trace_iters));
    if (rc) return rc;
    return out->status;
}

int so_register_batch(so_ctx* ctx, const void* surf, const uint32_t* n_points, size_t n_scans, size_t stride, size_t ioff,
                      const double* poses_in, const so_icp_opts* opts, so_icp_result* results) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !surf || !n_points || !poses_in || !opts || !results || stride < 12) return fail(SO_ERR_ARG, "bad args");
    if (n_scans == 0 || n_scans > c->max_batch) return fail(SO_ERR_CAPACITY, "n_scans exceeds so_config.max_batch");
    size_t total = 0;
    for (size_t s = 0; s < n_scans; ++s) { if (n_points[s] > c->cfg.max_scan_points) return fail(SO_ERR_CAPACITY, "scan too large"); total += n_points[s]; }
    SO_CUDA_TRY(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    int rc;
    if (stride == 16 && ioff == 12) {
        rc = register_core(c, c->d_scan, n_points, n_scans, poses_in, opts, results, false, surf);      // chunked, overlapped upload
    } else {
        rc = upload_cloud(c, surf, total, stride, ioff, c->d_scan);
        if (rc) return rc;
        rc = register_core(c, c->d_scan, n_points, n_scans, poses_in, opts, results, false);
    }
    if (rc) return rc;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (size_t s = 0; s < n_scans; ++s) results[s].time_total_ms = ms;
    return SO_OK;
}

int so_register_batch_edges(so_ctx* ctx, const void* surf, const uint32_t* n_points, const void* edge, const uint32_t* n_edge, size_t n_scans,
                            size_t stride, size_t ioff, const double* poses_in, const so_icp_opts* opts, so_icp_result* results) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !surf || !n_points || !n_edge || !poses_in || !opts || !results || stride < 12) return fail(SO_ERR_ARG, "bad args");
    if (n_scans == 0 || n_scans > c->max_batch) return fail(SO_ERR_CAPACITY, "n_scans exceeds so_config.max_batch");
    size_t total = 0, total_e = 0;
    for (size_t s = 0; s < n_scans; ++s) {
        if (n_points[s] > c->cfg.max_scan_points) return fail(SO_ERR_CAPACITY, "scan too large");
        total += n_points[s]; total_e += n_edge[s];
    }
    if (total_e > c->edge_cap) return fail(SO_ERR_CAPACITY, "edge clouds of the batch exceed so_config.max_scan_points points in total");
    if (total_e && !edge) return fail(SO_ERR_ARG, "bad args");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    int rc = upload_cloud(c, edge, total_e, stride, ioff, c->d_escan);
    if (rc) return rc;
    rc = upload_cloud(c, surf, total, stride, ioff, c->d_scan);
    if (rc) return rc;
    rc = register_core(c, c->d_scan, n_points, n_scans, poses_in, opts, results, false, nullptr, n_edge);
    if (rc) return rc;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (size_t s = 0; s < n_scans; ++s) results[s].time_total_ms = ms;
    return SO_OK;
}

int so_register_batch_device(so_ctx* ctx, const void* d_scans, const uint32_t* n_points, size_t n_scans, const double* poses_in,
                             const so_icp_opts* opts, so_icp_result* results) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !d_scans || !n_points || !poses_in || !opts || !results) return fail(SO_ERR_ARG, "bad args");
    if (n_scans == 0 || n_scans > c->max_batch) return fail(SO_ERR_CAPACITY, "n_scans exceeds so_config.max_batch");
    for (size_t s = 0; s < n_scans; ++s) if (n_points[s] > c->cfg.max_scan_points) return fail(SO_ERR_CAPACITY, "scan too large");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    int rc = register_core(c, static_cast<const float4*>(d_scans), n_points, n_scans, poses_in, opts, results, false);
    if (rc) return rc;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (size_t s = 0; s < n_scans; ++s) results[s].time_total_ms = ms;
    return SO_OK;
}

int so_set_pose_sink(so_ctx* ctx, void* d_rows, size_t cap_rows, size_t first_row) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || (d_rows && first_row > cap_rows)) return fail(SO_ERR_ARG, "bad args");
    // no synchronisation: rows are written by kernels on the context stream, so work already enqueued there keeps its order
    c->d_pose_sink = static_cast<double*>(d_rows); c->sink_cap = d_rows ? cap_rows : 0; c->sink_cursor = d_rows ? first_row : 0;
    return SO_OK;
}

// ---- stage-level entry points ------------------------------------------------------------------------------------
int so_correspond(so_ctx* ctx, const void* surf, size_t n, size_t stride, size_t ioff, const double pose[7], int32_t max_surface_features,
                  so_corr* corr, int32_t hist_obs[9], int32_t hist_rej[7]) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !surf || !pose || !corr || n == 0 || stride < 12) return fail(SO_ERR_ARG, "bad args");
    if (n > c->cfg.max_scan_points) return fail(SO_ERR_CAPACITY, "scan larger than so_config.max_scan_points");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    { int rc0 = ensure_maps(c); if (rc0) return rc0; }
    int rc = upload_cloud(c, surf, n, stride, ioff, c->d_scan);
    if (rc) return rc;
    so_icp_opts o{}; o.max_icp_iters = -1; o.lm_max_iterations = 4; o.max_surface_features = max_surface_features;
    init_state(c->h_state[0], pose, uint32_t(n), o);
    c->h_offset[0] = 0;
    SO_CUDA_TRY(cudaMemcpyAsync(c->d_state, c->h_state, sizeof(IcpState), cudaMemcpyHostToDevice, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(c->d_offset, c->h_offset, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    const uint32_t grid_x = (uint32_t(n) + kThreads - 1) / kThreads;
    const Chunk ch{0, 1, 0, uint32_t(n), grid_x};
    rc = prepare_scans(c, c->d_scan, ch, c->stream, false);              // stage API reports every point: decimation stays inside the kernels
    if (rc) return rc;
    const MapView mv = map_view(c, c->surf);
    BatchView bv = batch_view(c, c->d_scan_sorted);
    bv.coop_knn = grid_x <= kCoopMaxGrid && !c->no_coop_knn;            // as a single small registration would search
    timed_launch_begin(c); launch_correspond(mv, bv, c->corr, c->nn, grid_x, 1, c->stream); c->launches += kCorrLaunches + 2; timed_launch_end(c, 0);
    SO_CUDA_TRY(cudaGetLastError());
    std::vector<double4> nd(n); std::vector<double> w(n); std::vector<uchar4> fl(n); std::vector<uint32_t> nn(n * 5); std::vector<float> d2(n * 5);
    std::vector<float4> sorted(n);
    SO_CUDA_TRY(cudaMemcpyAsync(nd.data(), c->corr.nd, n * sizeof(double4), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(w.data(), c->corr.w, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(fl.data(), c->corr.flags, n * sizeof(uchar4), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(nn.data(), c->corr.nn, n * 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(d2.data(), c->corr.nn_d2, n * 5 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(sorted.data(), c->d_scan_sorted, n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(IcpState), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (size_t j = 0; j < n; ++j) {          // kernels work in cell order; report in the caller's order
        uint32_t orig;
        std::memcpy(&orig, &sorted[j].w, 4);
        so_corr& r = corr[orig];
        std::memset(&r, 0, sizeof(r));
        r.n[0] = nd[j].x; r.n[1] = nd[j].y; r.n[2] = nd[j].z; r.d = nd[j].w; r.w = w[j];
        for (int k = 0; k < 5; ++k) { r.nn[k] = nn[j * 5 + k]; r.nn_d2[k] = d2[j * 5 + k]; }
        r.status = fl[j].x; r.obs[0] = fl[j].y; r.obs[1] = fl[j].z; r.obs[2] = fl[j].w;
    }
    if (hist_obs) std::memcpy(hist_obs, c->h_state[0].hist_obs, 9 * sizeof(int32_t));
    if (hist_rej) std::memcpy(hist_rej, c->h_state[0].hist_rej, 7 * sizeof(int32_t));
    return SO_OK;
}

int so_correspond_edge(so_ctx* ctx, const void* edge, size_t n, size_t stride, size_t ioff, const double pose[7], so_edge_corr* corr,
                       int32_t hist_rej_line[7]) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !edge || !pose || !corr || n == 0 || stride < 12) return fail(SO_ERR_ARG, "bad args");
    if (n > c->edge_cap) return fail(SO_ERR_CAPACITY, "edge cloud larger than so_config.max_scan_points");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    { int rc0 = ensure_maps(c); if (rc0) return rc0; }
    int rc = upload_cloud(c, edge, n, stride, ioff, c->d_escan);
    if (rc) return rc;
    so_icp_opts o{}; o.max_icp_iters = -1; o.lm_max_iterations = 4;
    init_state(c->h_state[0], pose, 0, o, uint32_t(n));
    c->h_offset[0] = 0;
    SO_CUDA_TRY(cudaMemcpyAsync(c->d_state, c->h_state, sizeof(IcpState), cudaMemcpyHostToDevice, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(c->d_offset, c->h_offset, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    const uint32_t grid_e = (uint32_t(n) + kThreads - 1) / kThreads;
    const MapView me = map_view(c, c->edge);
    const BatchView bv = batch_view(c, c->d_scan_sorted);
    timed_launch_begin(c);
    launch_first_eval(bv, c->corr, 0, 1, c->stream, &me, &c->ebuf, grid_e);     // 0 plane CTAs: only the edge kernel + k_lm_step have work
    c->launches += 2;
    timed_launch_end(c, 4);
    SO_CUDA_TRY(cudaGetLastError());
    std::vector<double4> a(n), b(n); std::vector<uchar4> fl(n); std::vector<uint32_t> nn(n * 10), sel(n);
    SO_CUDA_TRY(cudaMemcpyAsync(a.data(), c->ebuf.a, n * sizeof(double4), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(b.data(), c->ebuf.b, n * sizeof(double4), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(fl.data(), c->ebuf.flags, n * sizeof(uchar4), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(nn.data(), c->ebuf.nn, n * 10 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(sel.data(), c->ebuf.selmask, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(IcpState), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; ++i) {
        so_edge_corr& r = corr[i];
        std::memset(&r, 0, sizeof(r));
        r.a[0] = a[i].x; r.a[1] = a[i].y; r.a[2] = a[i].z; r.w = a[i].w; r.b[0] = b[i].x; r.b[1] = b[i].y; r.b[2] = b[i].z;
        for (int k = 0; k < 10; ++k) r.nn[k] = nn[i * 10 + k];
        r.selected_mask = sel[i]; r.status = fl[i].x; r.n_selected = fl[i].y;
    }
    if (hist_rej_line) std::memcpy(hist_rej_line, c->h_state[0].hist_rej_line, 7 * sizeof(int32_t));
    return SO_OK;
}

int so_evaluate(so_ctx* ctx, const double pose[7], double H[36], double g[6], double* cost) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !pose) return fail(SO_ERR_ARG, "bad args");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    IcpState& s = c->h_state[0];
    const uint32_t n = uint32_t(s.n_points);
    const uint32_t grid_e = (uint32_t(s.n_edge) + kThreads - 1) / kThreads;
    if (n == 0 && grid_e == 0) return fail(SO_ERR_ARG, "so_evaluate needs a preceding so_correspond / so_correspond_edge");
    std::memcpy(s.cand, pose, 7 * sizeof(double));
    s.phase = PH_EVAL; s.max_icp_iters = -1;
    SO_CUDA_TRY(cudaMemcpyAsync(c->d_state, c->h_state, sizeof(IcpState), cudaMemcpyHostToDevice, c->stream));
    const uint32_t grid_x = (n + kThreads - 1) / kThreads;
    const BatchView bv = batch_view(c, c->d_scan_sorted);
    timed_launch_begin(c); launch_evaluate(bv, c->corr, grid_x, 1, c->stream, &c->ebuf, grid_e); c->launches++; timed_launch_end(c, 1);
    SO_CUDA_TRY(cudaGetLastError());
    SO_CUDA_TRY(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(IcpState), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (H) for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) H[i * 6 + j] = s.H[i <= j ? tri(i, j) : tri(j, i)];
    if (g) for (int i = 0; i < 6; ++i) g[i] = s.g[i];
    if (cost) *cost = s.cost;
    return SO_OK;
}

// ---- k-NN --------------------------------------------------------------------------------------------------------
int so_knn_device(so_ctx* ctx, const void* d_q, size_t nq, int k, float max_d2, uint32_t* d_idx, float* d_d2) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !d_q || !d_idx || !d_d2 || k < 1 || k > 8) return fail(SO_ERR_ARG, "bad args");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    { int rc0 = ensure_maps(c); if (rc0) return rc0; }
    if (nq == 0) return SO_OK;
    // Large query sets are first ordered by map cell (cell key -> radix sort); threads then take queries in that order and
    // write their answers at the caller's positions.  Small sets skip the ordering.
    const uint32_t* order = nullptr;
    const MapView mv = map_view(c, c->surf);
    timed_launch_begin(c);
    if (nq >= 32768 && nq < (size_t(1) << 31)) {
        int rc = query_sort_reserve(c, nq);
        if (rc) return rc;
        launch_query_keys(mv, static_cast<const float4*>(d_q), nq, c->d_qkeys, c->d_qvals, c->stream);
        rc = query_sort(c, nq);
        if (rc) return rc;
        order = c->d_qvals_out;
        c->launches++;
    }
    if (launch_knn(mv, static_cast<const float4*>(d_q), order, nq, k, max_d2, d_idx, d_d2, c->stream)) return fail(SO_ERR_ARG, "bad k");
    timed_launch_end(c, 2);
    SO_CUDA_TRY(cudaGetLastError());
    return SO_OK;
}

int so_knn(so_ctx* ctx, const float* q_xyz, size_t nq, size_t stride, int k, float max_d2, uint32_t* idx, float* d2) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !q_xyz || !idx || !d2 || k < 1 || k > 8 || stride < 12) return fail(SO_ERR_ARG, "bad args");
    SO_CUDA_TRY(cudaSetDevice(c->device));
    if (nq > c->knn_cap) {
        cudaFree(c->d_q); cudaFree(c->d_knn_idx); cudaFree(c->d_knn_d2);
        c->d_q = nullptr; c->d_knn_idx = nullptr; c->d_knn_d2 = nullptr; c->knn_cap = 0;
        SO_CUDA_TRY(cudaMalloc(&c->d_q, nq * sizeof(float4)));
        SO_CUDA_TRY(cudaMalloc(&c->d_knn_idx, nq * 8 * sizeof(uint32_t)));
        SO_CUDA_TRY(cudaMalloc(&c->d_knn_d2, nq * 8 * sizeof(float)));
        c->knn_cap = nq;
    }
    int rc = upload_cloud(c, q_xyz, nq, stride, stride /* no intensity */, c->d_q);
    if (rc) return rc;
    rc = so_knn_device(ctx, c->d_q, nq, k, max_d2, c->d_knn_idx, c->d_knn_d2);
    if (rc) return rc;
    SO_CUDA_TRY(cudaMemcpyAsync(idx, c->d_knn_idx, nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaMemcpyAsync(d2, c->d_knn_d2, nq * k * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SO_OK;
}

// ---- instrumentation -----------------------------------------------------------------------------------------------
uint64_t so_kernel_launches(so_ctx* ctx, int reset) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return 0;
    const uint64_t v = c->launches;
    if (reset) c->launches = 0;
    return v;
}

int so_sampling_indices(uint32_t n, int32_t max_surface_features, uint32_t* out, size_t cap, size_t* n_out) {
    if (!n_out || (!out && cap)) return fail(SO_ERR_ARG, "bad args");
    const double rate = (max_surface_features > 0 && n > uint32_t(max_surface_features)) ? 1.0 * max_surface_features / double(n) : -1.0;
    size_t k = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (rate >= 0.0 && !should_process_h(i, rate)) continue;
        if (k < cap) out[k] = i;
        ++k;
    }
    *n_out = k;
    return SO_OK;
}

int so_bytes_copied(so_ctx* ctx, uint64_t* h2d, uint64_t* d2h, int reset) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return fail(SO_ERR_ARG, "null ctx");
    if (h2d) *h2d = c->bytes_h2d;
    if (d2h) *d2h = c->bytes_d2h;
    if (reset) { c->bytes_h2d = 0; c->bytes_d2h = 0; }
    return SO_OK;
}

int so_build_flags(void) { return SO_FUSE_KNN_FIT ? SO_BUILD_FUSED_MATCH : 0; }

int so_profile_enable(so_ctx* ctx, int on) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return fail(SO_ERR_ARG, "null ctx");
    c->profiling = on != 0;
    return SO_OK;
}

int so_profile_get(so_ctx* ctx, int cls, double* ms, uint64_t* launches, int reset) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || cls < 0 || cls > 5) return fail(SO_ERR_ARG, "bad args");
    if (ms) *ms = c->prof[cls].ms;
    if (launches) *launches = c->prof[cls].launches;
    if (reset) c->prof[cls] = ProfileSlot{};
    return SO_OK;
}

}  // extern "C"
