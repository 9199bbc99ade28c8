// Internal declarations shared by the roft_b200 CUDA translation units (sm_100a).
// Not part of the public ABI (include/roft_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "roft_b200.h"

namespace roftb {

constexpr int kMaxFlows = ROFTB_MAX_DELAY;   // longest flow chain the FILTER LOOP warps a mask through
constexpr int kMaxChain = ROFTB_MAX_CHAIN;   // longest chain of the stand-alone operator (the stamped source's queue holds 30)
constexpr int kFrameRing = 16;               // frames the filter loop keeps addressable (> kMaxFlows)
constexpr int kFrameTable = 32;              // entries of the plane-pointer table passed to the kernels (> kMaxChain, >= kFrameRing)
constexpr int kThreads = 256;                // streaming kernels: 8 warps
constexpr int kWarpTilePx = 512;             // extract kernels: one warp covers 4 sub-tiles of 128 px
constexpr int kUnitPx = 128;                 // worklist granularity: 128 consecutive px = one quad per lane of a warp
constexpr int kBlockTilePx = kThreads / 32 * kWarpTilePx;  // 4096 px
// velocity kernel (velocity_track.cu): one cluster of up to kVtMaxCluster CTAs of kVtThreads threads per track
constexpr int kVtThreads = 256;
constexpr int kVtMaxCluster = 16;
constexpr int kVtMaxChunks = kVtMaxCluster * (kVtThreads / 32);   // one chunk of the unit list per warp
// per-CTA partials of the flow->velocity normal equations:
//   S1 = sum l L1^T L1 over index set {0,2,3,4,5} (15 upper-triangular entries)
//   S2 = sum l L2^T L2 over index set {1,2,3,4,5} (15)
//   g1 = sum l L1^T nu1 (5), g2 = sum l L2^T nu2 (5), count (1), min |n - m| (1)
constexpr int kVtPartN = 42;
constexpr int kAutoFp64Candidates = 32768;   // accum_fp64 == 2: tracks with fewer valid pixels accumulate in FP64
constexpr int kSelBins = 4096;               // radix-select histogram bins per level
// buffered velocities are capped at ROFTB_MAX_DELAY + 3 on the host: swap + (predict, correct) per buffered entry
constexpr int kMaxUkfBuffered = ROFTB_MAX_DELAY + 3;
constexpr int kMaxUkfOps = 2 * kMaxUkfBuffered + 2;

// ---- geometry / format of one batch of planes ----------------------------------------------
struct Geom {
    int W, H, HW;          // camera frame
    int Wf, Hf;            // flow frame
    int grid;              // W / Wf
    int flow_s16;          // 1: short2 elements, 0: float2
    float scale;           // flow scaling factor
    float inv_scale;       // 1 / scale
    int scale_mode;        // 0: scale == 1 (identity), 1: power of two (x * inv_scale is exact), 2: IEEE division
    float inv_grid;        // 1 / grid
    int grid_mode;         // same three cases for the division of the chased position by the grid size
    float cx, cy, inv_fx, inv_fy;
    double max_depth;      // depth gate (compared in double like the reference)
    float max_depth_f;     // smallest float >= max_depth: (double)d < max_depth  <=>  d < max_depth_f
    int stride;            // subsampling radius (>=1)
};

// table of the recent frames' device planes (passed to kernels by value)
struct FrameTable {
    const void* flow[kFrameTable];
    const float* depth[kFrameTable];
    long long flow_stride;   // scalar elements between tracks
    long long depth_stride;  // elements between tracks
};

// ---- mask synchronisation -----------------------------------------------------------------
struct WarpCtl {           // host-built, one per track per step
    int32_t has_new;       // a (stale) mask is delivered to this track at this frame
    int32_t first_mask;    // ... and it is the first one ever (initialisation, hpp:169-178)
    int32_t flow_valid;    // flow available && !is_first_frame (hpp:200-204)
    int32_t cur_slot;      // frame-ring slot of the current frame
    int32_t flow_aided;    // 0: plain segmentation source (no warp)
    int32_t reset;         // 1: forget the buffered flows first (source reset)
    int32_t pad[2];
};

struct MaskStat { int32_t nnz, vmin, vmax, pad; };

enum WarpMode : int32_t { kWarpCopyState = 0, kWarpCopyNew = 1, kWarpScatter = 2 };

struct WarpPlan {          // device-resolved by k_warp_plan
    int32_t mode;
    int32_t src_new;       // scatter source: 1 = newly delivered mask, 0 = state mask
    int32_t zero_origin;   // mask_(0,0) = 0 before findNonZero (hpp:224)
    int32_t n_flows;
    int32_t flow_slot[kMaxChain];
    int32_t uniform_val;   // >0: every non-zero source pixel has this value (byte-store fast path)
    int32_t dflt;          // value sampled by unmapped destinations = src(0,0) (Q2)
    int32_t fused;         // 1: the scatter of this track is done by the velocity pass (single current flow, state source)
    int32_t pad;
};

struct FlowBuf {           // per-track device mirror of flow_buffer_ (hpp:82,208,218)
    int32_t n;
    int32_t slot[kMaxFlows];
    int32_t uniform_val;   // single value of the current state mask (0 = mixed / unknown)
    int32_t pad[2];
};

// ---- velocity filter ----------------------------------------------------------------------
struct VelCtl {            // host-built, one per track per step
    int32_t enable;        // data_in: segmentation available, flow valid, not first frame
    int32_t prev_slot;     // frame-ring slot holding the previous depth
    int32_t cur_slot;      // slot of the current flow
    int32_t hist_slot;     // velocity-history ring slot to publish the twist to
    double dt;
};

// pool of scratch slots of the velocity kernel: a cluster claims one while it processes a track.  All of it is
// written and re-read within one kernel by the same cluster and overwritten in place by later tracks (L2-resident).
struct VelScratch {
    float2* nu;             // [S][cap] innovations (nu1, nu2) of the valid pixels, compacted per chunk, selection order
    uint2* dp;              // [S][cap] (depth bits, v << 16 | u) of the same pixels
    float* r;               // [S][cap] column-major pair norms r_i = |(nu[i], nu[N+i])| (SKFCorrection.cpp:93-94), compact
    long long cap;          // entries per slot (n_units * 128 + padding; a multiple of 4)
    uint32_t* hist;         // [S][3][kSelBins] radix-select histograms, one per level
    int32_t* chunk_cnt;     // [S][kVtMaxChunks] valid pixels per chunk
    double* part;           // [T][kVtMaxCluster][kVtPartN] per-CTA partial sums of pass B (per track: read by the epilogue kernel)
    double* track_sel;      // [T][4] (weights in use, b, valid pixels, CTAs) handed to the epilogue kernel
    double* sel_part;       // [S][kVtMaxCluster][4] per-CTA select statistics
    uint32_t* slot_bitmap;  // allocation bitmap (bits >= n_slots are pre-set)
    int n_slot_words;
    int n_slots;
    int32_t* track_slot;    // [T] slot claimed for a track (rank 0 -> the other CTAs)
    int32_t* chunk_aux;     // [T][kVtMaxChunks] candidate pixels per chunk (stride > 1)
};

// ---- pose UKF -----------------------------------------------------------------------------
// kOpCorrectBoth: ROFTFilter::correct_outlier_rejection (ROFTFilter.cpp:649-676) up to the choice: the standard correction
// and the velocity-only one (MeasurementMode::RepeatOnlyVelocity) of the same prediction are both computed and parked in
// UkfArgs::cand_*; the launch ends there for the track (UkfArgs::resume) and continues after the render-and-compare test.
enum UkfOpKind : int32_t { kOpPredict = 1, kOpCorrect = 2, kOpSwapBuffered = 3, kOpCorrectBoth = 4 };

struct UkfOp {
    int32_t kind;
    int32_t meas_type;     // ROFTB_MEAS_* for kOpCorrect
    int32_t vel_slot;      // velocity-history slot providing the (v, w) part, -1: use meas[0..5]
    int32_t pad;
    double dt;             // kOpPredict
    double meas[13];       // (v, w, x, q) - pose part filled by the host
};

struct UkfParams {
    double alpha, beta, kappa;
    double psd_lin[3], sigma_ang[3];
    double cov_v[3], cov_w[3], cov_x[3], cov_q[3];
};

// ---- launchers (defined in the .cu files) ---------------------------------------------------
struct MaskSyncArgs {
    Geom g;
    FrameTable ft;
    int n_tracks;
    const uint8_t* new_mask; long long new_stride;   // may be null
    const uint8_t* state_src; uint8_t* state_dst;    // [T][HW]
    int32_t* winner;                                  // [T][HW]
    const WarpCtl* ctl;                               // device
    MaskStat* stat; WarpPlan* plan; FlowBuf* fbuf;    // device
    int segm_delay;
    // worklists (launch_tile_list) of the state mask and of the newly delivered mask
    const int32_t* s_list; const int32_t* s_n;
    const int32_t* n_list; const int32_t* n_n;
    int n_warp_tiles;      // 128-px units per plane
    int fuse;              // allow k_warp_plan to hand single-flow propagation to the velocity pass
    // occupancy flags [T][n_units] of the two state planes (may be null in operator mode)
    const uint8_t* occ_src; uint8_t* occ_dst;
    unsigned long long* span_clock;   // diagnostics (may be null): [first start, last end] of init / scatter / gather
};
// the mask synchronisation in two halves: (stats, plan, init) must precede a fused velocity pass; (scatter of the
// non-fused tracks, gather) may run concurrently with the velocity passes
// stats of a new mask (unless the caller gathered them with launch_tile_list) + the per-track plan
int launch_mask_plan(const MaskSyncArgs& a, cudaStream_t s, bool have_stats = false);
int launch_mask_init(const MaskSyncArgs& a, cudaStream_t s);   // destination plane / winner plane / flags of the non-fused tracks
int launch_mask_plan_init(const MaskSyncArgs& a, cudaStream_t s, bool planned = false);
int launch_mask_scatter_gather(const MaskSyncArgs& a, cudaStream_t s);
// planned = true: a.plan was filled by the caller (operator mode), skip the stats / plan kernels
int launch_mask_sync(const MaskSyncArgs& a, cudaStream_t s, bool planned = false);
int launch_threshold(const uint8_t* src, uint8_t* dst, size_t n, cudaStream_t s);

struct VelocityArgs {
    Geom g;
    FrameTable ft;
    int n_tracks;
    const uint8_t* seg; long long seg_stride; int thr;   // selected iff byte > thr
    const uint8_t* occ_src;                               // [T][n_units] occupancy flags of seg (1: the unit holds a non-zero byte)
    const VelCtl* ctl;                                    // device
    int weight_flow;
    VelScratch scratch;
    // state
    double* v_mean; double* v_cov;     // [T][6], [T][36]
    const double* q_diag;              // [6] process noise (device)
    double r_flow[2];
    double fx, fy, cx, cy;
    int accum_fp64;                    // pass B per-pixel terms and sums in FP64 (else FP32 terms)
    double* vel_hist; int hist_ring;   // [T][ring][6]
    // diagnostics
    int32_t* out_count; double* out_lambda; double* out_eta;  // device [T], [T][36], [T][6]
    int32_t* wl_units; int32_t* wl_pixels;                    // device [T]: listed units, candidate pixels (may be null)
    // largest-first scheduling (may be null): clusters take the tracks in `order`; the last cluster to finish writes
    // `order_next` from wl_units for the following step; done_ticket is a zero-initialised counter
    const int32_t* order; int32_t* order_next; uint32_t* done_ticket;
    // optional: two more streams (+ fork / join events) so that the biggest tracks run in larger clusters beside the rest
    cudaStream_t side_stream[2]; cudaEvent_t side_fork; cudaEvent_t side_join[2];
    unsigned long long* phase_clock;                          // device [T][8] globaltimer stamps at the phase boundaries (may be null)
    unsigned long long* span_clock;                           // diagnostics (may be null): [first start, last end] of the launch
    // fused mask propagation (see WarpPlan::fused): destination plane and its occupancy flags
    int fuse_scatter; const WarpPlan* plan; uint8_t* state_dst; uint8_t* occ_dst;
    int update_state;                  // 0: only compute lambda/eta/count (operator mode)
    const double* x_pred_override;     // operator mode: [T][6] predicted mean for the innovations (else v_mean)
};
int launch_velocity(const VelocityArgs& a, cudaStream_t s);
// per-device set-up of the velocity kernel (shared-memory opt-in, cluster attributes); returns the number of clusters
// that can be resident at once = the number of scratch slots worth allocating
int velocity_prepare_device(const Geom& g, int n_units, int* max_active_clusters);
int velocity_cluster_size();
// occupancy flags of a plane that comes from outside, and the list of flagged units
int launch_unit_flags(const uint8_t* plane, long long stride, int HW, int n_items, uint8_t* flags, cudaStream_t s);
int launch_flag_list(const uint8_t* flags, int n_units, int n_items, int32_t* wt_list, int32_t* wt_n, const WarpPlan* plan,
                     cudaStream_t s);
int launch_order_from_flags(const uint8_t* flags, int n_units, int n_tracks, int32_t* units_tmp, uint32_t* ticket, int32_t* order,
                            cudaStream_t s);
// worklist of the non-empty 128-px units of a byte plane (+ rank base of the bytes > thr); active: optional per-item
// flags, item i is processed iff active[i * active_stride] != 0
// stat (optional): non-zero count / min / max of every processed plane, accumulated in the same read
int launch_tile_list(const uint8_t* plane, long long stride, int thr, int HW, int n_items, int32_t* wt_count, int32_t* wt_list,
                     int32_t* wt_n, const int32_t* active, int active_stride, cudaStream_t s, MaskStat* stat = nullptr);
// per-warp-tile exclusive prefix of the number of pixels with byte > thr (row-major rank base); ctl may be null
int launch_mask_rank(const uint8_t* seg, long long seg_stride, int thr, int HW, int n_items, int32_t* wt_count, int32_t* total,
                     const VelCtl* ctl, cudaStream_t s);
int launch_wt_scan(int32_t* wt_count, int n_warp_tiles, int n_items, int32_t* total, cudaStream_t s);

struct UkfArgs {
    int n_tracks;
    UkfParams p;
    const UkfOp* ops; const int32_t* n_ops; int max_ops;   // device [T][max_ops], [T]
    double* mean; double* cov;                              // [T][13], [T][144]
    double* buf_mean; double* buf_cov;                      // buffered belief (may be null)
    const double* vel_hist; int hist_ring;                  // [T][ring][6] (may be null)
    unsigned long long* span_clock;                         // diagnostics (may be null): [first start, last end]
    // outlier rejection (all null when off): per-track op index to start from (-1: nothing left), candidates [T][2][13|144]
    int32_t* resume; double* cand_mean; double* cand_cov;
    double* warm;   // [T][2][145] eigenvectors of the last covariance square roots (+ use counter), may be null: cold Jacobi
};
int launch_ukf(const UkfArgs& a, cudaStream_t s);
int ukf_prepare_device();

// depth rasteriser (render.cu): n_items poses of ONE mesh -> n_items tiles of w x h pixels
struct RenderArgs {
    int n_items;
    const float* vertices; int n_vertices;   // device [n_vertices][3], model frame, metres
    const int32_t* faces; int n_faces;       // device [n_faces][3]
    const float* model;                      // device [n_items][12]: rotation row-major (9), translation (3)
    float fx, fy, cx, cy;                    // intrinsics of the tile (camera intrinsics / divider)
    int w, h;
    const float* scale; int n_scale;         // optional device [n_scale][3]: vertex scale of item i = scale[i % n_scale]
};
size_t render_vertex_scratch_bytes(int n_items, int n_vertices);
int launch_render_depth(const RenderArgs& a, void* vertex_scratch, uint32_t* zbuf, float* out, cudaStream_t s);
int launch_pick_best(int n, const double* err, const int32_t* samples, double gain, int32_t* selected, double* likelihoods,
                     cudaStream_t s);
int launch_or_features(const Geom& g, int n_tracks, const UkfOp* ops, int max_ops, int bits, const uint8_t* mask, long long mask_stride,
                       int thr, const float* depth, long long depth_stride, int32_t* wt_count, int32_t* total, uint2* feat,
                       long long feat_stride, int32_t* n_feat, cudaStream_t st);
int launch_or_feat_copy(int n_tracks, const UkfOp* ops, int max_ops, int bits, const uint2* src, const int32_t* n_src, uint2* dst,
                        int32_t* n_dst, long long stride, cudaStream_t st);
int launch_or_l1(int n_tracks, const int32_t* resume, const uint2* feat, const int32_t* n_feat, long long feat_stride,
                 const float* rendered, long long tile, int divider, int W, double* err, int32_t* samples, cudaStream_t st);
int launch_or_models(int n_tracks, const int32_t* resume, const double* cand_mean, float* model, cudaStream_t s);
int launch_or_select(int n_tracks, const int32_t* resume, const int32_t* selected, const double* cand_mean, const double* cand_cov,
                     double* mean, double* cov, cudaStream_t s);

struct SelectArgs {        // ordered compaction helpers (export / points / L1)
    Geom g;
    int n_items;
    const uint8_t* mask; long long mask_stride; int thr;
    const float* depth; long long depth_stride;
    const void* flow; long long flow_stride;
    int32_t* wt_count;     // [n][n_warp_tiles]
    int32_t* wt_count2;    // [n][n_warp_tiles]
};
int launch_export_measurement(const SelectArgs& a, double dt, double fx, double fy, double cx, double cy, int capacity,
                              double* z, double* H, int32_t* n_valid, cudaStream_t s);
int launch_masked_points(const SelectArgs& a, double max_depth, double fx, double fy, double cx, double cy, int capacity,
                         double* points, int32_t* count, cudaStream_t s);
int launch_masked_depth_l1(const SelectArgs& a, const float* rendered, long long rendered_stride, int divider,
                           double* err_sum, int32_t* samples, cudaStream_t s);

// number of kernel launches issued through ROFTB_LAUNCH (for bench.py "gpu_launches")
extern std::atomic<long long> g_launch_count;

}  // namespace roftb

#define ROFTB_LAUNCH(kernel, grid, block, smem, stream, ...)            \
    do {                                                                \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);     \
        ++::roftb::g_launch_count;                                      \
    } while (0)

// ---- device helpers ---------------------------------------------------------------------------
namespace roftb {

// C `int(float)` as compiled for x86-64 (cvttss2si): truncation toward zero, NaN / overflow -> INT_MIN.
__device__ __forceinline__ int cvt_int(float t) {
    return (fabsf(t) < 2147483648.0f) ? (int)t : (int)0x80000000;
}

__device__ __forceinline__ uint32_t ld_nc_u32(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_nc_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// x / scale in IEEE FP32 (ImageOpticalFlowMeasurement.hpp:249-250): identity for scale 1, an exact multiplication
// for a power-of-two scale (NVOF's 32), a correctly rounded division otherwise - bit-identical in all three cases
__device__ __forceinline__ float div_scale(float x, const Geom& g) {
    return g.scale_mode == 0 ? x : g.scale_mode == 1 ? x * g.inv_scale : __fdiv_rn(x, g.scale);
}
__device__ __forceinline__ float div_grid(float x, const Geom& g) {
    return g.grid_mode == 0 ? x : g.grid_mode == 1 ? x * g.inv_grid : __fdiv_rn(x, (float)g.grid);
}

// one flow element (dx, dy) = float(f) / scale in FP32
__device__ __forceinline__ float2 load_flow(const void* base, long long idx, const Geom& g) {
    float2 f;
    if (g.flow_s16) {
        short2 v = __ldg(reinterpret_cast<const short2*>(base) + idx);
        f = make_float2((float)v.x, (float)v.y);
    } else {
        f = __ldg(reinterpret_cast<const float2*>(base) + idx);
    }
    f.x = div_scale(f.x, g);
    f.y = div_scale(f.y, g);
    return f;
}

// OpticalFlowUtilities.h:19-22 (a NaN fails the magnitude test, so the isnan tests are implied)
__device__ __forceinline__ bool flow_valid(float dx, float dy) { return fabsf(dx) < 1e9f && fabsf(dy) < 1e9f; }

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// diagnostics: first start / last end of a kernel's blocks on the global timer (both words preset to all ones)
__device__ __forceinline__ void span_stamp(unsigned long long* span, bool end) {
    if (!span || threadIdx.x != 0) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMin(span + (end ? 1 : 0), end ? ~t : t);  // (the end is kept complemented: one preset, all ones, for both)
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// inclusive scan across the warp
__device__ __forceinline__ int warp_scan_incl(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

}  // namespace roftb
