// C ABI of libroft_b200.so (include/roft_b200.h): context, the batched ROFTFilter loop and the
// stateless operators.  Host-side logic here is the part of the reference that is pure control flow:
//   ROFTFilter::filtering_step sequencing                 src/roft-lib/src/ROFTFilter.cpp:255-367
//   ImageSegmentationOFAidedSource::step_frame (the content-independent part)   ...OFAidedSource.hpp:128-231
//   ImageOpticalFlowMeasurement::freeze early-outs        ...ImageOpticalFlowMeasurement.hpp:184-229
//   CartesianQuaternionMeasurement::freeze mode machine   src/roft-lib/src/CartesianQuaternionMeasurement.cpp:92-348
// Everything value-dependent (empty masks, observability, the filters) runs on the device, so a step
// never synchronises with the GPU.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "roftb_internal.cuh"

using namespace roftb;

namespace {

std::string g_create_error;
constexpr int kCtlRing = 16;   // in-flight per-step control blocks
constexpr int kHistRing = 32;  // velocity history ring (> ROFTB_MAX_DELAY + 3 + kUkfLag)
constexpr int kUkfLag = 10;    // steps the pose UKF stream may trail the streaming kernels (absorbs the re-sync replay burst)

struct PoseMeasHost {          // CartesianQuaternionMeasurement state (.h:98-140), values replaced by slots
    std::deque<int> buffer;    // buffer_velocities_ as velocity-history slots
    bool is_pose = false;
    bool is_first_velocity_in = false;
    int last_vel_slot = -1;
    double last_pose[7] = {0, 0, 0, 1, 0, 0, 0};
    int mtype = ROFTB_MEAS_NONE;
    int meas_vel_slot = -1;    // velocity part of the current measurement_
};

struct TrackHost {
    bool seg_src_available = false;   // ImageSegmentationOFAidedSource::segmentation_available_
    bool of_first_frame = true;       // ImageSegmentationOFAidedSource::is_first_frame_
    bool segmeas_available = false;   // ImageSegmentationMeasurement::segmentation_available_
    bool fm_first_frame = true;       // ImageOpticalFlowMeasurement::is_first_frame_
    int prev_slot = -1;               // frame slot of previous_depth_
    bool or_features_initialized = false;  // ROFTFilter::outlier_rejection_features_initialized_
    PoseMeasHost pm;
};

}  // namespace

// Pipelined parts: a context of many tracks is run as two (to four) complete sub-contexts (own streams, events, control
// rings, state) over consecutive pieces of the track range.  Tracks never interact, so the halves are independent
// pipelines - and the tail of one half's velocity launch, its epilogue and the launch latencies of its next step overlap
// the other half's bulk instead of leaving the GPU partly idle (16 % of a step at 256 tracks).  The second
// half's host work (control blocks, ~25 CUDA calls per step) runs on a worker thread.
constexpr int kMaxParts = 4;
struct PartWorker {  // one host thread per extra part: runs that part's share of an API call
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> task;
    bool has_task = false, done = false, quit = false;
    int result = 0;
    void post(const std::function<int()>& f) {
        {
            std::lock_guard<std::mutex> lk(m);
            task = f;
            has_task = true;
            done = false;
        }
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return done; });
        return result;
    }
    void loop() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return has_task || quit; });
            if (quit) break;
            std::function<int()> t = std::move(task);
            has_task = false;
            lk.unlock();
            const int r = t();
            lk.lock();
            result = r;
            done = true;
            cv.notify_all();
        }
    }
    void stop() {
        if (!th.joinable()) return;
        {
            std::lock_guard<std::mutex> lk(m);
            quit = true;
        }
        cv.notify_all();
        th.join();
    }
};
struct Composite {
    int n = 0;                                   // parts
    roftb_ctx* sub[kMaxParts] = {nullptr, nullptr, nullptr, nullptr};
    int first[kMaxParts] = {0, 0, 0, 0}, count[kMaxParts] = {0, 0, 0, 0};
    cudaEvent_t join_ev[kMaxParts] = {nullptr, nullptr, nullptr, nullptr};
    PartWorker worker[kMaxParts];                // [0] unused: part 0 runs on the caller's thread
    // f(part) for every part at once
    int run_all(const std::function<int(int)>& f) {
        for (int h = 1; h < n; ++h) worker[h].post([&f, h] { return f(h); });
        int rc = f(0);
        for (int h = 1; h < n; ++h) {
            const int r = worker[h].wait();
            if (!rc) rc = r;
        }
        return rc;
    }
};

struct roftb_ctx {
    roftb_config cfg;
    Composite* comp = nullptr;   // non-null: this context only fans out to two half contexts
    Geom g;
    int T = 0;
    size_t HW = 0;
    size_t flow_elems = 0;  // scalar elements per track
    int dev = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, ukf_stream = nullptr, mask_stream = nullptr, prep_stream = nullptr,
                 aux_stream = nullptr;
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    // render-and-compare pose outlier rejection inside the filter loop (cfg.outlier_rejection)
    struct OutlierRejection {
        int divider = 4;
        int32_t* resume = nullptr;          // [T] op index the UKF continues from after the test (-1: nothing pending)
        double* cand_mean = nullptr;        // [T][2][13]
        double* cand_cov = nullptr;         // [T][2][144]
        float* model = nullptr;             // [2][T][12]
        void* vertex_scratch = nullptr;     // (allocated with the mesh)
        uint32_t* zbuf = nullptr;           // [2][T][h][w]
        float* rendered = nullptr;          // [2][T][h][w]
        double* err = nullptr; int32_t* samples = nullptr;   // [2][T]
        int32_t* selected = nullptr;        // [T]
        int32_t* wt_count = nullptr;        // [T][n_units] rank scratch of the masked L1
        // features buffered at the previous re-synchronisation (ROFTFilter.cpp:624-646), compacted: per track the list
        // (pixel index, depth) of every second segmentation pixel (extract.cu: k_or_features).  A refresh goes through a
        // staging list: stage <- live planes as soon as the step's mask state is final (own stream; the planes are recycled
        // two steps later), snapshot <- stage on the pose stream, before or after the step's test, which may still be
        // reading the previous snapshot
        uint2* feat_snap = nullptr; uint2* feat_stage = nullptr;     // [T][feat_stride]
        int32_t* n_snap = nullptr; int32_t* n_stage = nullptr;       // [T] entries
        int32_t* rank_total = nullptr;                               // [T] scratch
        long long feat_stride = 0;
        cudaEvent_t stage_done[3] = {nullptr, nullptr, nullptr}, unstage_done = nullptr, ops_ready = nullptr;
        bool stage_done_used[3] = {false, false, false}, unstage_done_used = false;
        cudaStream_t stream = nullptr;      // snapshot copies
    } orj;
    double* ukf_warm = nullptr;                      // [T][2][145] warm start of the UKF's Jacobi decompositions
    float* mesh_vertices = nullptr;                  // outlier-rejection mesh (roftb_set_mesh): device [n][3]
    int32_t* mesh_faces = nullptr;
    int mesh_nv = 0, mesh_nf = 0;
    float* mesh_scale = nullptr;                     // optional [T][3] (roftb_set_mesh_scale)
    unsigned long long* span_clock = nullptr;        // diagnostics: [8 steps][init, scatter, gather, velocity, ukf][2]
    cudaStream_t vel_side[2] = {nullptr, nullptr};   // larger-cluster launches of the velocity kernel (biggest tracks)
    cudaEvent_t vel_fork = nullptr, vel_join[2] = {nullptr, nullptr};
    cudaEvent_t prep_event[2] = {nullptr, nullptr}, prep2_event[2] = {nullptr, nullptr}, vel_done_event = nullptr;
    bool vel_done_event_used = false;
    cudaEvent_t vel_event[kCtlRing], ukf_event[kCtlRing], join_event = nullptr, plan_event = nullptr, mask_event = nullptr;
    bool ukf_event_used[kCtlRing];
    bool mask_event_used = false;
    std::string err;
    long long launches0 = 0;
    bool poisoned = false;  // a hard error inside roftb_filter_step left the host state machines ahead of the device

    // device state
    uint8_t* mask_state[3] = {nullptr, nullptr, nullptr};  // ring of 3: step k reads [k%3], writes [(k+1)%3]
    uint8_t* mask_occ[3] = {nullptr, nullptr, nullptr};    // occupancy flags of the three planes: [T][n_units]
    int mask_cur = 0;
    int32_t* winner = nullptr;
    VelScratch scratch;             // scratch-slot pool of the velocity kernel
    int32_t* wl_units = nullptr;    // [T] listed units / candidate pixels of the last step (diagnostics)
    int32_t* wl_pixels = nullptr;
    unsigned long long* phase_clock = nullptr;  // [T][8] phase stamps of the velocity kernel (profiling)
    int32_t* vel_order = nullptr;   // [2][T] scheduling order of the velocity kernel (by step parity), largest worklist first
    uint32_t* vel_ticket = nullptr;
    int32_t* order_units = nullptr;   // flagged units per track of the freshly synchronised state (k_order_from_flags)
    int32_t* wt_list = nullptr;     // state-mask worklist (only for masks scattered by the stand-alone kernel)
    int32_t* wt_n = nullptr;
    int32_t* nl_count = nullptr;    // newly delivered mask worklist
    int32_t* nl_list = nullptr;
    int32_t* nl_n = nullptr;
    int32_t* wt_count2 = nullptr;
    int n_warp_tiles = 0;   // 512-px tiles (extract kernels)
    int n_units = 0;        // 128-px worklist units
    MaskStat* stat = nullptr;
    WarpPlan* plan = nullptr;
    FlowBuf* fbuf = nullptr;
    double *v_mean = nullptr, *v_cov = nullptr, *p_mean = nullptr, *p_cov = nullptr, *pb_mean = nullptr, *pb_cov = nullptr;
    double* vel_hist = nullptr;
    double* q_diag = nullptr;
    int32_t* d_count = nullptr;
    double *d_lambda = nullptr, *d_eta = nullptr;
    UkfParams ukf_p;

    // per-step control blocks: pinned host ring + one device copy
    WarpCtl* h_wctl = nullptr; VelCtl* h_vctl = nullptr; UkfOp* h_ops = nullptr; int32_t* h_nops = nullptr;
    WarpCtl* d_wctl = nullptr; VelCtl* d_vctl = nullptr; UkfOp* d_ops = nullptr; int32_t* d_nops = nullptr;
    cudaEvent_t ctl_event[kCtlRing];
    bool ctl_event_used[kCtlRing];

    // frames
    FrameTable ft;
    long long frame_idx = 0;
    std::vector<TrackHost> th;
    // host-path staging ring (lazily allocated)
    float* stage_depth = nullptr; void* stage_flow = nullptr; uint8_t* stage_mask = nullptr;
    cudaEvent_t copy_done = nullptr, stage_event = nullptr;
    uint8_t* thr_tmp = nullptr;

    // optional per-phase device timing (bench.py roofline): 8 events per in-flight step
    bool prof_on = false;
    cudaEvent_t prof_ev[kCtlRing][11];
    bool prof_used[kCtlRing];
    double prof_ms[7];
    double prof_phase[4];   // time of the velocity kernel's phases summed over tracks (ns): A, select, B, epilogue
    long long prof_steps = 0;
};

namespace {

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);               \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

int fail(roftb_ctx* ctx, const char* msg) {
    ctx->err = msg;
    return -2;
}

template <class T>
cudaError_t dalloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, n * sizeof(T));
    return e;
}

size_t flow_scalar_bytes(const roftb_ctx* ctx) { return ctx->g.flow_s16 ? 2 : 4; }

int check_geom(const roftb_config& c, std::string& err) {
    if (c.n_tracks <= 0) { err = "n_tracks must be > 0"; return -2; }
    if (c.width <= 0 || c.height <= 0 || (c.width % 4) != 0 || ((long long)c.width * c.height) % 16 != 0) {
        err = "width must be a multiple of 4 and width*height a multiple of 16"; return -2;
    }
    if (c.flow_format != ROFTB_FLOW_F32 && c.flow_format != ROFTB_FLOW_S16) { err = "unknown flow_format"; return -2; }
    if (c.flow_grid <= 0 || c.width / c.flow_grid <= 0 || c.height / c.flow_grid <= 0) { err = "bad flow_grid"; return -2; }
    if (!(c.flow_scale > 0.f)) { err = "flow_scale must be > 0"; return -2; }
    if (c.subsampling_radius < 1) { err = "subsampling_radius must be >= 1"; return -2; }
    if (c.segm_delay > ROFTB_MAX_DELAY || c.pose_delay > ROFTB_MAX_DELAY) { err = "delay exceeds ROFTB_MAX_DELAY"; return -2; }
    if (!(c.fx > 0) || !(c.fy > 0)) { err = "fx, fy must be > 0"; return -2; }
    return 0;
}

void fill_geom(roftb_ctx* ctx) {
    const roftb_config& c = ctx->cfg;
    Geom& g = ctx->g;
    g.W = c.width; g.H = c.height; g.HW = c.width * c.height;
    g.grid = c.flow_grid;
    // DatasetImageOpticalFlow.cpp:46: grid = width / cols  =>  cols = ceil-free width / grid for exact multiples
    g.Wf = c.width / c.flow_grid; g.Hf = c.height / c.flow_grid;
    g.flow_s16 = c.flow_format == ROFTB_FLOW_S16;
    g.scale = c.flow_scale;
    g.inv_scale = 1.0f / c.flow_scale;
    auto mode_of = [](float v) {
        if (v == 1.0f) return 0;
        int e = 0;
        return std::frexp(v, &e) == 0.5f ? 1 : 2;  // power of two
    };
    g.scale_mode = mode_of(c.flow_scale);
    g.inv_grid = 1.0f / (float)c.flow_grid;
    g.grid_mode = mode_of((float)c.flow_grid);
    g.cx = (float)c.cx; g.cy = (float)c.cy;
    g.inv_fx = (float)(1.0 / c.fx); g.inv_fy = (float)(1.0 / c.fy);
    g.max_depth = c.depth_maximum;
    g.max_depth_f = (float)c.depth_maximum;
    if ((double)g.max_depth_f < c.depth_maximum) g.max_depth_f = std::nextafterf(g.max_depth_f, INFINITY);
    g.stride = c.subsampling_radius;
}

}  // namespace

extern "C" {

int roftb_version(void) { return ROFTB_VERSION; }

void roftb_config_default(roftb_config* c) {
    // config/config_fast_ycb.cfg
    memset(c, 0, sizeof(*c));
    c->n_tracks = 1;
    c->width = 1280; c->height = 720;
    c->fx = 1229.4285612615463; c->fy = 1229.4285612615463; c->cx = 640.0; c->cy = 360.0;
    c->sample_time = 0.033333333333;
    c->flow_format = ROFTB_FLOW_F32; c->flow_grid = 1; c->flow_scale = 1.0f;
    c->cov_flow[0] = c->cov_flow[1] = 1.0;
    c->depth_maximum = 2.0;
    c->subsampling_radius = 35;
    c->weight_flow = 1;
    for (int i = 0; i < 6; ++i) { c->v_sigma[i] = 0.1; c->v_cov0[i] = 1e-3; }
    for (int i = 0; i < 3; ++i) {
        c->p_sigma_linear[i] = 1.0; c->p_sigma_angular[i] = 1.0;
        c->cov_v[i] = 0.1; c->cov_w[i] = 1e-4; c->cov_x[i] = 1e-3; c->cov_q[i] = 1e-4;
    }
    for (int i = 0; i < 12; ++i) c->p_cov0[i] = 1e-3;
    c->ut_alpha = 1.0; c->ut_beta = 2.0; c->ut_kappa = 0.0;
    c->use_pose = 1; c->use_pose_resync = 1; c->use_velocity = 1; c->flow_aided = 1;
    c->segm_delay = 6; c->pose_delay = 6;
    c->device = 0;
    c->accum_fp64 = 2;
    c->outlier_rejection = 0;          // (cfg:110 enables it; here it needs roftb_set_mesh first)
    c->outlier_rejection_divider = 0;
    c->outlier_rejection_gain = 1.0;   // 0.01 in the cfg, `true` in ROFTFilter (see roft_b200.h)
}

const char* roftb_last_error(const roftb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int64_t roftb_kernel_launches(const roftb_ctx* ctx) { return ctx ? (int64_t)(g_launch_count.load() - ctx->launches0) : 0; }

void* roftb_stream(roftb_ctx* ctx) { return ctx ? (void*)(ctx->comp ? ctx->comp->sub[0]->stream : ctx->stream) : nullptr; }

static int create_single(const roftb_config* cfg, roftb_ctx** out) {
    if (!cfg || !out) { g_create_error = "null argument"; return -2; }
    *out = nullptr;
    std::string err;
    if (check_geom(*cfg, err)) { g_create_error = err; return -2; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                         " (roft_b200 has no CPU fallback)";
        return -1;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "bad device ordinal"; return -2; }
    roftb_ctx* ctx = new roftb_ctx();
    ctx->cfg = *cfg;
    ctx->T = cfg->n_tracks;
    ctx->dev = cfg->device;
    fill_geom(ctx);
    ctx->HW = (size_t)ctx->g.HW;
    ctx->flow_elems = (size_t)ctx->g.Wf * ctx->g.Hf * 2;
    ctx->launches0 = g_launch_count;
    for (int i = 0; i < kCtlRing; ++i) {
        ctx->ctl_event_used[i] = false;
        ctx->prof_used[i] = false;
        ctx->ukf_event_used[i] = false;
        ctx->vel_event[i] = nullptr;
        ctx->ukf_event[i] = nullptr;
        for (int j = 0; j < 11; ++j) ctx->prof_ev[i][j] = nullptr;
    }
    for (int j = 0; j < 7; ++j) ctx->prof_ms[j] = 0.0;
    memset(&ctx->ft, 0, sizeof(ctx->ft));
    ctx->th.assign(ctx->T, TrackHost());
    const int T = ctx->T;
    const size_t HW = ctx->HW;
    ctx->n_warp_tiles = (int)((HW + kWarpTilePx - 1) / kWarpTilePx);
    ctx->n_units = (int)((HW + kUnitPx - 1) / kUnitPx);
    const int n_block_tiles = (int)((HW + kBlockTilePx - 1) / kBlockTilePx);
    (void)n_block_tiles;
    memset(&ctx->scratch, 0, sizeof(ctx->scratch));
#define CKC(call)                                                              \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) {                                              \
            g_create_error = std::string(#call) + ": " + cudaGetErrorString(e__); \
            roftb_destroy(ctx);                                                \
            return -1;                                                         \
        }                                                                      \
    } while (0)
    CKC(cudaSetDevice(ctx->dev));
    // Stream priorities: all equal by default.  Measured on B200: the velocity kernel keeps blocks
    // pending for most of a step, and whichever way the other streams are ranked around it (above: their blocks displace
    // its clusters; below: they wait for its tail) the step takes as long or longer - the work of an event step
    // (new-mask scatter, pose re-sync replay) has to be paid in machine time either way.
    // ROFTB_STREAM_PRIORITIES=1: scatter / preparation / UKF above the velocity kernel (diagnostic).
    int prio_lo = 0, prio_hi = 0;
    CKC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));  // numerically lower = higher priority
    const char* env_prio = getenv("ROFTB_STREAM_PRIORITIES");
    const bool flat = !(env_prio && env_prio[0] == '1');
    const bool side_first = env_prio && env_prio[0] == '2';  // only the larger-cluster launches above everything else
    const int prio_main = (flat && !side_first) ? prio_hi : std::min(prio_lo, prio_hi + 1);
    const int prio_side = side_first ? prio_hi : prio_main;
    if (side_first) prio_hi = prio_main;
    const int prio_ukf = prio_hi;
    CKC(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_main));
    CKC(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, prio_main));
    CKC(cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, prio_hi));
    CKC(cudaStreamCreateWithPriority(&ctx->ukf_stream, cudaStreamNonBlocking, prio_ukf));
    CKC(cudaEventCreateWithFlags(&ctx->join_event, cudaEventDisableTiming));
    CKC(cudaStreamCreateWithPriority(&ctx->mask_stream, cudaStreamNonBlocking, prio_hi));
    CKC(cudaStreamCreateWithPriority(&ctx->prep_stream, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < 2; ++i) {
        CKC(cudaStreamCreateWithPriority(&ctx->vel_side[i], cudaStreamNonBlocking, prio_side));
        CKC(cudaEventCreateWithFlags(&ctx->vel_join[i], cudaEventDisableTiming));
    }
    CKC(cudaEventCreateWithFlags(&ctx->vel_fork, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->prep_event[0], cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->prep_event[1], cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->prep2_event[0], cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->prep2_event[1], cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->vel_done_event, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->plan_event, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->mask_event, cudaEventDisableTiming));
    for (int i = 0; i < kCtlRing; ++i) {
        CKC(cudaEventCreateWithFlags(&ctx->vel_event[i], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&ctx->ukf_event[i], cudaEventDisableTiming));
    }
    CKC(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->stage_event, cudaEventDisableTiming));
    for (int i = 0; i < kCtlRing; ++i) CKC(cudaEventCreateWithFlags(&ctx->ctl_event[i], cudaEventDisableTiming));
    for (int i = 0; i < kCtlRing; ++i)
        for (int j = 0; j < 11; ++j) CKC(cudaEventCreate(&ctx->prof_ev[i][j]));
    CKC(dalloc(&ctx->mask_state[0], T * HW));
    CKC(dalloc(&ctx->mask_state[1], T * HW));
    CKC(dalloc(&ctx->mask_state[2], T * HW));
    CKC(dalloc(&ctx->winner, T * HW));
    for (int i = 0; i < 3; ++i) CKC(dalloc(&ctx->mask_occ[i], (size_t)T * ctx->n_units));
    {
        // velocity kernel: shared-memory / cluster attributes are per device; one scratch slot per cluster that can
        // be resident at once
        int max_clusters = 0;
        if (ukf_prepare_device()) {
            g_create_error = "cannot configure the UKF kernel";
            roftb_destroy(ctx);
            return -1;
        }
        if (velocity_prepare_device(ctx->g, ctx->n_units, &max_clusters)) {
            g_create_error = std::string("velocity kernel set-up failed: ") + cudaGetErrorString(cudaGetLastError());
            roftb_destroy(ctx);
            return -1;
        }
        VelScratch& sc = ctx->scratch;
        sc.n_slots = std::max(1, std::min(T, max_clusters));
        sc.n_slot_words = (sc.n_slots + 31) / 32;
        sc.cap = (long long)ctx->n_units * kUnitPx + 64;
        const size_t S = (size_t)sc.n_slots;
        CKC(dalloc(&sc.nu, S * (size_t)sc.cap));
        CKC(dalloc(&sc.dp, S * (size_t)sc.cap));
        CKC(dalloc(&sc.r, S * (size_t)sc.cap));
        CKC(dalloc(&sc.hist, S * 3 * kSelBins));
        CKC(dalloc(&sc.chunk_cnt, S * kVtMaxChunks));
        CKC(dalloc(&sc.part, (size_t)T * kVtMaxCluster * kVtPartN));
        CKC(dalloc(&sc.track_sel, (size_t)T * 4));
        CKC(dalloc(&sc.sel_part, S * kVtMaxCluster * 4));
        CKC(dalloc(&sc.slot_bitmap, (size_t)sc.n_slot_words));
        CKC(dalloc(&sc.track_slot, (size_t)T));
        CKC(dalloc(&sc.chunk_aux, (size_t)T * kVtMaxChunks));
    }
    CKC(dalloc(&ctx->wl_units, (size_t)T));
    CKC(dalloc(&ctx->wl_pixels, (size_t)T));
    CKC(dalloc(&ctx->phase_clock, (size_t)T * 8));
    CKC(dalloc(&ctx->span_clock, (size_t)8 * 10));
    CKC(cudaMemset(ctx->span_clock, 0xFF, 8 * 10 * sizeof(unsigned long long)));
    CKC(dalloc(&ctx->vel_order, (size_t)2 * T));
    CKC(dalloc(&ctx->vel_ticket, (size_t)2));  // [0] velocity kernel, [1] k_order_from_flags
    CKC(dalloc(&ctx->order_units, (size_t)T));
    CKC(dalloc(&ctx->wt_count2, (size_t)T * ctx->n_warp_tiles));
    CKC(dalloc(&ctx->wt_list, (size_t)2 * T * ctx->n_units));   // x2: worklists / plans are double-buffered by step parity
    CKC(dalloc(&ctx->wt_n, (size_t)4 * T));
    CKC(dalloc(&ctx->nl_count, (size_t)2 * T * ctx->n_units));
    CKC(dalloc(&ctx->nl_list, (size_t)2 * T * ctx->n_units));
    CKC(dalloc(&ctx->nl_n, (size_t)4 * T));
    CKC(dalloc(&ctx->stat, (size_t)T));
    CKC(dalloc(&ctx->plan, (size_t)2 * T));
    CKC(dalloc(&ctx->fbuf, (size_t)T));
    CKC(dalloc(&ctx->v_mean, (size_t)T * 6));
    CKC(dalloc(&ctx->v_cov, (size_t)T * 36));
    CKC(dalloc(&ctx->p_mean, (size_t)T * 13));
    CKC(dalloc(&ctx->p_cov, (size_t)T * 144));
    CKC(dalloc(&ctx->pb_mean, (size_t)T * 13));
    CKC(dalloc(&ctx->pb_cov, (size_t)T * 144));
    CKC(dalloc(&ctx->ukf_warm, (size_t)T * 290));
    CKC(dalloc(&ctx->vel_hist, (size_t)T * kHistRing * 6));
    if (cfg->outlier_rejection) {
        auto& oj = ctx->orj;
        oj.divider = cfg->outlier_rejection_divider > 0 ? cfg->outlier_rejection_divider : (cfg->width == 640 ? 2 : 4);
        if (cfg->width % oj.divider || cfg->height % oj.divider) {
            g_create_error = "outlier_rejection_divider must divide the frame size";
            roftb_destroy(ctx);
            return -1;
        }
        const size_t tile = (size_t)(cfg->width / oj.divider) * (cfg->height / oj.divider);
        CKC(dalloc(&oj.resume, (size_t)T));
        CKC(dalloc(&oj.cand_mean, (size_t)T * 26));
        CKC(dalloc(&oj.cand_cov, (size_t)T * 288));
        CKC(dalloc(&oj.model, (size_t)T * 24));
        CKC(dalloc(&oj.zbuf, (size_t)2 * T * tile));
        CKC(dalloc(&oj.rendered, (size_t)2 * T * tile));
        CKC(dalloc(&oj.err, (size_t)2 * T));
        CKC(dalloc(&oj.samples, (size_t)2 * T));
        CKC(dalloc(&oj.selected, (size_t)T));
        CKC(dalloc(&oj.wt_count, (size_t)T * ctx->n_units));
        oj.feat_stride = (long long)((ctx->HW / 2 + 16) & ~(size_t)1);
        CKC(dalloc(&oj.feat_snap, (size_t)T * oj.feat_stride));
        CKC(dalloc(&oj.feat_stage, (size_t)T * oj.feat_stride));
        CKC(dalloc(&oj.n_snap, (size_t)T));
        CKC(dalloc(&oj.n_stage, (size_t)T));
        CKC(dalloc(&oj.rank_total, (size_t)T));
        for (int i = 0; i < 3; ++i) CKC(cudaEventCreateWithFlags(&oj.stage_done[i], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&oj.unstage_done, cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&oj.ops_ready, cudaEventDisableTiming));
        CKC(cudaStreamCreateWithFlags(&oj.stream, cudaStreamNonBlocking));
    }
    CKC(dalloc(&ctx->q_diag, (size_t)6));
    CKC(dalloc(&ctx->d_count, (size_t)T));
    CKC(dalloc(&ctx->d_lambda, (size_t)T * 36));
    CKC(dalloc(&ctx->d_eta, (size_t)T * 6));
    CKC(dalloc(&ctx->d_wctl, (size_t)2 * T));
    CKC(dalloc(&ctx->d_vctl, (size_t)T));
    CKC(dalloc(&ctx->d_ops, (size_t)T * kMaxUkfOps * kCtlRing));
    CKC(dalloc(&ctx->d_nops, (size_t)T * kCtlRing));
    CKC(cudaMallocHost(&ctx->h_wctl, sizeof(WarpCtl) * T * kCtlRing));
    CKC(cudaMallocHost(&ctx->h_vctl, sizeof(VelCtl) * T * kCtlRing));
    CKC(cudaMallocHost(&ctx->h_ops, sizeof(UkfOp) * T * kMaxUkfOps * kCtlRing));
    CKC(cudaMallocHost(&ctx->h_nops, sizeof(int32_t) * T * kCtlRing));
    {
        std::vector<MaskStat> st(T, MaskStat{0, 255, 0, 0});
        CKC(cudaMemcpy(ctx->stat, st.data(), sizeof(MaskStat) * T, cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(ctx->q_diag, cfg->v_sigma, sizeof(double) * 6, cudaMemcpyHostToDevice));
    }
    UkfParams& p = ctx->ukf_p;
    p.alpha = cfg->ut_alpha; p.beta = cfg->ut_beta; p.kappa = cfg->ut_kappa;
    for (int i = 0; i < 3; ++i) {
        p.psd_lin[i] = cfg->p_sigma_linear[i]; p.sigma_ang[i] = cfg->p_sigma_angular[i];
        p.cov_v[i] = cfg->cov_v[i]; p.cov_w[i] = cfg->cov_w[i]; p.cov_x[i] = cfg->cov_x[i]; p.cov_q[i] = cfg->cov_q[i];
    }
#undef CKC
    if (roftb_filter_init(ctx, nullptr, nullptr)) {
        g_create_error = ctx->err;
        roftb_destroy(ctx);
        return -1;
    }
    *out = ctx;
    return 0;
}

int roftb_create(const roftb_config* cfg, roftb_ctx** out) {
    if (!cfg || !out) { g_create_error = "null argument"; return -2; }
    // ROFTB_PARTS: number of pipelined part contexts (1 = none); default: 2 from 64 tracks
    const char* eh = getenv("ROFTB_PARTS");
    int parts = eh ? atoi(eh) : (cfg->n_tracks >= 64 ? 2 : 1);
    parts = std::max(1, std::min(parts, std::min(kMaxParts, cfg->n_tracks)));
    if (parts < 2) return create_single(cfg, out);
    *out = nullptr;
    roftb_ctx* ctx = new roftb_ctx();
    ctx->cfg = *cfg;
    ctx->T = cfg->n_tracks;
    ctx->dev = cfg->device;
    ctx->launches0 = g_launch_count;
    ctx->comp = new Composite();
    Composite& c = *ctx->comp;
    c.n = parts;
    for (int h = 0, at = 0; h < parts; ++h) {
        c.first[h] = at;
        c.count[h] = cfg->n_tracks / parts + (h < cfg->n_tracks % parts ? 1 : 0);
        at += c.count[h];
    }
    for (int h = 0; h < parts; ++h) {
        roftb_config sc = *cfg;
        sc.n_tracks = c.count[h];
        if (create_single(&sc, &c.sub[h]) != 0) {
            roftb_destroy(ctx);
            return -1;
        }
    }
    fill_geom(ctx);
    ctx->HW = (size_t)ctx->g.HW;
    ctx->flow_elems = (size_t)ctx->g.Wf * ctx->g.Hf * 2;
    if (cudaSetDevice(ctx->dev) != cudaSuccess) {
        g_create_error = "cudaSetDevice failed";
        roftb_destroy(ctx);
        return -1;
    }
    for (int h = 1; h < parts; ++h) {
        if (cudaEventCreateWithFlags(&c.join_ev[h], cudaEventDisableTiming) != cudaSuccess) {
            g_create_error = "cannot create the join events of the part contexts";
            roftb_destroy(ctx);
            return -1;
        }
        PartWorker* w = &c.worker[h];
        w->th = std::thread([w] { w->loop(); });
    }
    *out = ctx;
    return 0;
}

// error of a half context -> the composite's
static int comp_fail(roftb_ctx* ctx, int rc) {
    if (rc) {
        for (int h = 0; h < ctx->comp->n; ++h)
            if (ctx->comp->sub[h] && !ctx->comp->sub[h]->err.empty()) { ctx->err = ctx->comp->sub[h]->err; break; }
    }
    return rc;
}

void roftb_destroy(roftb_ctx* ctx) {
    if (!ctx) return;
    if (ctx->comp) {
        Composite* c = ctx->comp;
        for (int h = 1; h < kMaxParts; ++h) c->worker[h].stop();
        for (int h = 0; h < kMaxParts; ++h) {
            roftb_destroy(c->sub[h]);
            if (c->join_ev[h]) cudaEventDestroy(c->join_ev[h]);
        }
        delete c;
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->dev);
    // nothing may still be running on any of the streams when the buffers go away
    for (cudaStream_t st : {ctx->copy_stream, ctx->prep_stream, ctx->stream, ctx->aux_stream, ctx->mask_stream, ctx->ukf_stream,
                            ctx->vel_side[0], ctx->vel_side[1]})
        if (st) cudaStreamSynchronize(st);
    for (int i = 0; i < 2; ++i) {
        if (ctx->vel_side[i]) cudaStreamDestroy(ctx->vel_side[i]);
        if (ctx->vel_join[i]) cudaEventDestroy(ctx->vel_join[i]);
    }
    if (ctx->vel_fork) cudaEventDestroy(ctx->vel_fork);
    {
        auto& oj = ctx->orj;
        if (oj.stream) { cudaStreamSynchronize(oj.stream); cudaStreamDestroy(oj.stream); }
        void* op[] = {oj.resume, oj.cand_mean, oj.cand_cov, oj.model, oj.vertex_scratch, oj.zbuf, oj.rendered, oj.err, oj.samples,
                      oj.selected, oj.wt_count, oj.feat_snap, oj.feat_stage, oj.n_snap, oj.n_stage, oj.rank_total};
        for (void* p : op)
            if (p) cudaFree(p);
        for (int i = 0; i < 3; ++i)
            if (oj.stage_done[i]) cudaEventDestroy(oj.stage_done[i]);
        if (oj.unstage_done) cudaEventDestroy(oj.unstage_done);
        if (oj.ops_ready) cudaEventDestroy(oj.ops_ready);
    }
    void* dptrs[] = {ctx->mask_state[0], ctx->mask_state[1], ctx->mask_state[2], ctx->mask_occ[0], ctx->mask_occ[1], ctx->mask_occ[2],
                     ctx->winner, ctx->scratch.nu, ctx->scratch.dp, ctx->scratch.r, ctx->scratch.hist, ctx->scratch.chunk_cnt,
                     ctx->scratch.part, ctx->scratch.track_sel, ctx->scratch.sel_part, ctx->scratch.slot_bitmap, ctx->scratch.track_slot, ctx->scratch.chunk_aux,
                     ctx->wl_units, ctx->wl_pixels, ctx->phase_clock, ctx->span_clock, ctx->mesh_vertices, ctx->mesh_faces, ctx->mesh_scale, ctx->ukf_warm, ctx->vel_order, ctx->vel_ticket, ctx->order_units,
                     ctx->wt_count2, ctx->wt_list, ctx->wt_n, ctx->nl_count, ctx->nl_list, ctx->nl_n, ctx->stat, ctx->plan, ctx->fbuf, ctx->v_mean,
                     ctx->v_cov, ctx->p_mean, ctx->p_cov, ctx->pb_mean, ctx->pb_cov, ctx->vel_hist, ctx->q_diag,
                     ctx->d_count, ctx->d_lambda, ctx->d_eta, ctx->d_wctl, ctx->d_vctl, ctx->d_ops, ctx->d_nops,
                     ctx->stage_depth, ctx->stage_flow, ctx->stage_mask, ctx->thr_tmp};
    for (void* p : dptrs)
        if (p) cudaFree(p);
    if (ctx->h_wctl) cudaFreeHost(ctx->h_wctl);
    if (ctx->h_vctl) cudaFreeHost(ctx->h_vctl);
    if (ctx->h_ops) cudaFreeHost(ctx->h_ops);
    if (ctx->h_nops) cudaFreeHost(ctx->h_nops);
    for (int i = 0; i < kCtlRing; ++i)
        if (ctx->ctl_event[i]) cudaEventDestroy(ctx->ctl_event[i]);
    for (int i = 0; i < kCtlRing; ++i)
        for (int j = 0; j < 11; ++j)
            if (ctx->prof_ev[i][j]) cudaEventDestroy(ctx->prof_ev[i][j]);
    for (int i = 0; i < kCtlRing; ++i) {
        if (ctx->vel_event[i]) cudaEventDestroy(ctx->vel_event[i]);
        if (ctx->ukf_event[i]) cudaEventDestroy(ctx->ukf_event[i]);
    }
    if (ctx->join_event) cudaEventDestroy(ctx->join_event);
    if (ctx->plan_event) cudaEventDestroy(ctx->plan_event);
    if (ctx->mask_event) cudaEventDestroy(ctx->mask_event);
    if (ctx->mask_stream) { cudaStreamSynchronize(ctx->mask_stream); cudaStreamDestroy(ctx->mask_stream); }
    if (ctx->prep_stream) { cudaStreamSynchronize(ctx->prep_stream); cudaStreamDestroy(ctx->prep_stream); }
    if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
    if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
    if (ctx->aux_join) cudaEventDestroy(ctx->aux_join);
    for (int i = 0; i < 2; ++i) {
        if (ctx->prep_event[i]) cudaEventDestroy(ctx->prep_event[i]);
        if (ctx->prep2_event[i]) cudaEventDestroy(ctx->prep2_event[i]);
    }
    if (ctx->vel_done_event) cudaEventDestroy(ctx->vel_done_event);
    if (ctx->ukf_stream) { cudaStreamSynchronize(ctx->ukf_stream); cudaStreamDestroy(ctx->ukf_stream); }
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    if (ctx->stage_event) cudaEventDestroy(ctx->stage_event);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int roftb_sync(roftb_ctx* ctx) {
    if (!ctx) return -2;
    if (ctx->comp) {
        int rc = 0;
        for (int h = 0; h < ctx->comp->n; ++h) rc |= roftb_sync(ctx->comp->sub[h]);
        return comp_fail(ctx, rc);
    }
    CK(cudaSetDevice(ctx->dev));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->prep_stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->mask_stream));
    CK(cudaStreamSynchronize(ctx->ukf_stream));
    return 0;
}

int roftb_join(roftb_ctx* ctx) {
    if (!ctx) return -2;
    if (ctx->comp) {  // everything of EVERY part before whatever is recorded on roftb_stream (= the first part's) next
        Composite& c = *ctx->comp;
        for (int h = c.n - 1; h >= 0; --h)
            if (roftb_join(c.sub[h])) return comp_fail(ctx, -1);
        CK(cudaSetDevice(ctx->dev));
        for (int h = 1; h < c.n; ++h) {
            CK(cudaEventRecord(c.join_ev[h], c.sub[h]->stream));
            CK(cudaStreamWaitEvent(c.sub[0]->stream, c.join_ev[h], 0));
        }
        return 0;
    }
    CK(cudaSetDevice(ctx->dev));
    CK(cudaEventRecord(ctx->join_event, ctx->ukf_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->join_event, 0));
    if (ctx->mask_event_used) CK(cudaStreamWaitEvent(ctx->stream, ctx->mask_event, 0));
    return 0;
}

static void prof_collect(roftb_ctx* ctx, int slot) {
    if (!ctx->prof_used[slot]) return;
    cudaEventSynchronize(ctx->prof_ev[slot][2]);
    cudaEventSynchronize(ctx->prof_ev[slot][7]);
    cudaEventSynchronize(ctx->prof_ev[slot][9]);
    // event pairs: prep stream (10 -> 0), velocity kernel (1 -> 2), new-mask scatter (6 -> 7), pose UKF (8 -> 9)
    static const int kE0[4] = {10, 1, 6, 8}, kE1[4] = {0, 2, 7, 9}, kDst[4] = {0, 1, 5, 6};
    for (int j = 0; j < 4; ++j) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->prof_ev[slot][kE0[j]], ctx->prof_ev[slot][kE1[j]]) == cudaSuccess) ctx->prof_ms[kDst[j]] += ms;
    }
    if (getenv("ROFTB_TIMELINE")) {
        // start / end of the step's kernels on their streams, relative to the first profiled step (diagnostic)
        static cudaEvent_t base = nullptr;
        if (!base) base = ctx->prof_ev[slot][1];
        float t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        static const int ev[8] = {10, 0, 1, 2, 6, 7, 8, 9};
        for (int j = 0; j < 8; ++j) cudaEventElapsedTime(&t[j], base, ctx->prof_ev[slot][ev[j]]);
        fprintf(stderr, "[timeline] slot %2d  init+list %.3f-%.3f | velocity %.3f-%.3f | scatter %.3f-%.3f | ukf %.3f-%.3f\n", slot,
                t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7]);
    }
    ctx->prof_steps++;
    ctx->prof_used[slot] = false;
}

int roftb_profile(roftb_ctx* ctx, int32_t enable, double* ms_per_step, int64_t* steps) {
    if (!ctx) return -2;
    if (ctx->comp) {  // phase times: mean over the parts (they run side by side)
        Composite& c = *ctx->comp;
        double acc[7] = {0, 0, 0, 0, 0, 0, 0};
        int64_t s0 = 0;
        int rc = 0;
        for (int h = 0; h < c.n; ++h) {
            double m[7] = {0, 0, 0, 0, 0, 0, 0};
            int64_t sh = 0;
            rc |= roftb_profile(c.sub[h], enable, m, &sh);
            for (int j = 0; j < 7; ++j) acc[j] += m[j] / c.n;
            if (h == 0) s0 = sh;
        }
        if (ms_per_step)
            for (int j = 0; j < 7; ++j) ms_per_step[j] = acc[j];
        if (steps) *steps = s0;
        return comp_fail(ctx, rc);
    }
    CK(cudaSetDevice(ctx->dev));
    for (int i = 0; i < kCtlRing; ++i) prof_collect(ctx, i);
    if (ms_per_step) {
        // the velocity kernel is ONE launch: its device time (prof_ms[1]) is split over pass A / select / pass B /
        // epilogue in proportion to the per-track phase stamps of the last profiled step
        double ph[4] = {0, 0, 0, 0};
        if (ctx->prof_steps) {
            CK(cudaStreamSynchronize(ctx->stream));
            std::vector<unsigned long long> clk((size_t)ctx->T * 8);
            CK(cudaMemcpy(clk.data(), ctx->phase_clock, clk.size() * 8, cudaMemcpyDeviceToHost));
            for (int t = 0; t < ctx->T; ++t) {
                const unsigned long long* c = &clk[(size_t)t * 8];
                if (!c[0] || c[7] < c[0]) continue;  // track took no part (or only propagated)
                ph[0] += (double)(c[2] - c[0]);
                ph[1] += (double)(c[5] >= c[2] ? c[5] - c[2] : 0);
                ph[2] += (double)(c[5] >= c[2] ? c[6] - c[5] : c[6] - c[2]);
                ph[3] += (double)(c[7] - c[6]);
            }
        }
        if (getenv("ROFTB_PHASE_DEBUG") && ctx->prof_steps) {
            // mean per-track latency of each interval between the phase stamps (ns) and the concurrency available
            double iv[7] = {0, 0, 0, 0, 0, 0, 0};
            int nt = 0;
            std::vector<unsigned long long> clk((size_t)ctx->T * 8);
            cudaMemcpy(clk.data(), ctx->phase_clock, clk.size() * 8, cudaMemcpyDeviceToHost);
            unsigned long long first = ~0ull, last = 0;
            for (int t = 0; t < ctx->T; ++t) {
                const unsigned long long* c = &clk[(size_t)t * 8];
                if (!c[0] || c[7] < c[0]) continue;
                ++nt;
                for (int j = 0; j < 7; ++j) iv[j] += (c[j + 1] >= c[j] && c[j]) ? (double)(c[j + 1] - c[j]) : 0.0;
                first = std::min(first, c[0]);
                last = std::max(last, c[7]);
            }
            {
                std::vector<int32_t> wu(ctx->T);
                cudaMemcpy(wu.data(), ctx->wl_units, wu.size() * 4, cudaMemcpyDeviceToHost);
                double tmax = 0, tsum = 0;
                int umax = 0, umin = 1 << 30, t_of_max = 0;
                long long usum = 0;
                for (int t = 0; t < ctx->T; ++t) {
                    const unsigned long long* c = &clk[(size_t)t * 8];
                    if (!c[0] || c[7] < c[0]) continue;
                    const double d = (double)(c[7] - c[0]);
                    tsum += d;
                    if (d > tmax) { tmax = d; t_of_max = t; }
                    umax = std::max(umax, wu[t]); umin = std::min(umin, wu[t]); usum += wu[t];
                }
                fprintf(stderr, "[roftb] units per track min %d mean %.0f max %d; track latency mean %.1f us max %.1f us (track %d, %d units)\n",
                        umin, nt ? (double)usum / nt : 0.0, umax, nt ? tsum / nt * 1e-3 : 0.0, tmax * 1e-3, t_of_max, wu[t_of_max]);
            }
            if (atoi(getenv("ROFTB_PHASE_DEBUG")) >= 2 && nt) {  // clusters in flight over the launch, start/end of the biggest tracks
                std::vector<int32_t> wu(ctx->T);
                cudaMemcpy(wu.data(), ctx->wl_units, wu.size() * 4, cudaMemcpyDeviceToHost);
                const double span = (double)(last - first);
                fprintf(stderr, "[roftb] clusters in flight at 5%% steps of the span:");
                for (int q = 0; q < 20; ++q) {
                    const unsigned long long at = first + (unsigned long long)(span * (q + 0.5) / 20.0);
                    int n_in = 0;
                    for (int t = 0; t < ctx->T; ++t) {
                        const unsigned long long* c = &clk[(size_t)t * 8];
                        if (c[0] && c[0] <= at && c[7] > at) ++n_in;
                    }
                    fprintf(stderr, " %d", n_in);
                }
                fprintf(stderr, "\n");
                std::vector<int> idx(ctx->T);
                for (int t = 0; t < ctx->T; ++t) idx[t] = t;
                std::sort(idx.begin(), idx.end(), [&](int x, int y) { return wu[x] > wu[y]; });
                for (int i = 0; i < std::min(ctx->T, 6); ++i) {
                    const unsigned long long* c = &clk[(size_t)idx[i] * 8];
                    fprintf(stderr, "[roftb]   track %d (%d units): start +%.1f us, end +%.1f us\n", idx[i], wu[idx[i]],
                            (double)(c[0] - first) * 1e-3, (double)(c[7] - first) * 1e-3);
                }
            }
            if (atoi(getenv("ROFTB_PHASE_DEBUG")) >= 2) {  // first start / last end of the kernels of the last 8 steps
                cudaDeviceSynchronize();
                unsigned long long sc[80];
                cudaMemcpy(sc, ctx->span_clock, sizeof(sc), cudaMemcpyDeviceToHost);
                unsigned long long t0 = ~0ull;
                for (int i = 0; i < 80; i += 2) t0 = std::min(t0, sc[i]);
                static const char* kn[5] = {"init", "scatter", "gather", "velocity", "ukf"};
                for (int q = 0; q < 8; ++q) {
                    const long long fi = ctx->frame_idx - 8 + q;
                    if (fi < 0) continue;
                    const unsigned long long* c = sc + (size_t)(fi % 8) * 10;
                    fprintf(stderr, "[roftb] step %lld:", fi);
                    for (int j = 0; j < 5; ++j)
                        if (c[2 * j] != ~0ull) fprintf(stderr, " %s %.0f-%.0f", kn[j], (double)(c[2 * j] - t0) * 1e-3, (double)(~c[2 * j + 1] - t0) * 1e-3);
                    fprintf(stderr, "\n");
                }
            }
            fprintf(stderr, "[roftb] velocity kernel: cluster %d, slots %d, tracks %d, kernel span %.1f us; per-track mean ns:"
                            " prologue %.0f | passA %.0f | pair %.0f | level1 %.0f | level2 %.0f | passB %.0f | epilogue %.0f\n",
                    velocity_cluster_size(), ctx->scratch.n_slots, nt, nt ? (double)(last - first) * 1e-3 : 0.0,
                    iv[0] / std::max(nt, 1), iv[1] / std::max(nt, 1), iv[2] / std::max(nt, 1), iv[3] / std::max(nt, 1),
                    iv[4] / std::max(nt, 1), iv[5] / std::max(nt, 1), iv[6] / std::max(nt, 1));
        }
        const double tot = ph[0] + ph[1] + ph[2] + ph[3];
        const double n = ctx->prof_steps ? (double)ctx->prof_steps : 1.0;
        const double vel = ctx->prof_ms[1] / n;
        ms_per_step[0] = ctx->prof_ms[0] / n;
        for (int j = 0; j < 4; ++j) ms_per_step[1 + j] = tot > 0 ? vel * ph[j] / tot : (j == 0 ? vel : 0.0);
        ms_per_step[5] = ctx->prof_ms[5] / n;
        ms_per_step[6] = ctx->prof_ms[6] / n;
    }
    if (steps) *steps = ctx->prof_steps;
    if (enable != (ctx->prof_on ? 1 : 0) || enable) {
        for (int j = 0; j < 7; ++j) ctx->prof_ms[j] = 0.0;
        ctx->prof_steps = 0;
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaMemset(ctx->phase_clock, 0, (size_t)ctx->T * 8 * sizeof(unsigned long long)));
    }
    ctx->prof_on = enable != 0;
    return 0;
}

// ---------------------------------------------------------------------------------------------
int roftb_filter_init(roftb_ctx* ctx, const double* p_mean0, const double* v_mean0) {
    if (!ctx) return -2;
    if (ctx->comp) {
        Composite& c = *ctx->comp;
        int rc = 0;
        for (int h = 0; h < c.n; ++h)
            rc |= roftb_filter_init(c.sub[h], p_mean0 ? p_mean0 + (size_t)c.first[h] * 13 : nullptr, v_mean0 ? v_mean0 + (size_t)c.first[h] * 6 : nullptr);
        return comp_fail(ctx, rc);
    }
    const int T = ctx->T;
    CK(cudaSetDevice(ctx->dev));
    CK(cudaStreamSynchronize(ctx->prep_stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->mask_stream));
    CK(cudaStreamSynchronize(ctx->ukf_stream));
    ctx->mask_event_used = false;
    ctx->vel_done_event_used = false;
    if (ctx->orj.stream) CK(cudaStreamSynchronize(ctx->orj.stream));
    for (int i = 0; i < 3; ++i) ctx->orj.stage_done_used[i] = false;
    ctx->orj.unstage_done_used = false;
    // ROFTFilter::initialization_step (ROFTFilter.cpp:216-237)
    std::vector<double> pm((size_t)T * 13, 0.0), pc((size_t)T * 144, 0.0), vm((size_t)T * 6, 0.0), vc((size_t)T * 36, 0.0);
    for (int t = 0; t < T; ++t) {
        if (p_mean0) memcpy(&pm[(size_t)t * 13], p_mean0 + (size_t)t * 13, sizeof(double) * 13);
        else pm[(size_t)t * 13 + 9] = 1.0;
        if (v_mean0) memcpy(&vm[(size_t)t * 6], v_mean0 + (size_t)t * 6, sizeof(double) * 6);
        for (int i = 0; i < 12; ++i) pc[(size_t)t * 144 + i * 13] = ctx->cfg.p_cov0[i];
        for (int i = 0; i < 6; ++i) vc[(size_t)t * 36 + i * 7] = ctx->cfg.v_cov0[i];
    }
    CK(cudaMemcpy(ctx->p_mean, pm.data(), pm.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->p_cov, pc.data(), pc.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->pb_mean, pm.data(), pm.size() * 8, cudaMemcpyHostToDevice));  // buffered_belief_ = p_corr_belief_
    CK(cudaMemcpy(ctx->pb_cov, pc.data(), pc.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->v_mean, vm.data(), vm.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->v_cov, vc.data(), vc.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(ctx->vel_hist, 0, sizeof(double) * T * kHistRing * 6));
    CK(cudaMemset(ctx->ukf_warm, 0, sizeof(double) * T * 290));  // use counters 0: the first decompositions start cold
    CK(cudaMemset(ctx->mask_state[0], 0, (size_t)T * ctx->HW));
    CK(cudaMemset(ctx->mask_state[1], 0, (size_t)T * ctx->HW));
    CK(cudaMemset(ctx->mask_state[2], 0, (size_t)T * ctx->HW));
    CK(cudaMemset(ctx->fbuf, 0, sizeof(FlowBuf) * T));
    for (int i = 0; i < 3; ++i) CK(cudaMemset(ctx->mask_occ[i], 0, (size_t)T * ctx->n_units));
    {
        // every scratch slot free; the bits past the last slot stay set so they are never handed out
        std::vector<uint32_t> bm(ctx->scratch.n_slot_words, 0u);
        for (int b = ctx->scratch.n_slots; b < ctx->scratch.n_slot_words * 32; ++b) bm[b >> 5] |= 1u << (b & 31);
        CK(cudaMemcpy(ctx->scratch.slot_bitmap, bm.data(), bm.size() * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMemset(ctx->wl_units, 0, sizeof(int32_t) * T));
    CK(cudaMemset(ctx->wl_pixels, 0, sizeof(int32_t) * T));
    {
        std::vector<int32_t> ord((size_t)2 * T);
        for (int t = 0; t < T; ++t) ord[t] = ord[(size_t)T + t] = t;
        CK(cudaMemcpy(ctx->vel_order, ord.data(), ord.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(ctx->vel_ticket, 0, 8));
    }
    CK(cudaMemset(ctx->d_count, 0, sizeof(int32_t) * T));
    ctx->mask_cur = 0;
    ctx->frame_idx = 0;
    ctx->poisoned = false;
    ctx->th.assign(T, TrackHost());  // segmentation_->reset() etc.
    return 0;
}

// Stage host planes of this frame into the device ring; returns the device pointers.
static int stage_host_frame(roftb_ctx* ctx, const roftb_frame* f, int slot, const float** d_depth, const void** d_flow,
                            const uint8_t** d_mask) {
    const size_t T = ctx->T, HW = ctx->HW;
    const size_t fb = flow_scalar_bytes(ctx);
    if (!ctx->stage_depth) {
        CK(cudaMalloc(&ctx->stage_depth, sizeof(float) * T * HW * kFrameRing));
        CK(cudaMalloc(&ctx->stage_flow, fb * T * ctx->flow_elems * kFrameRing));
        CK(cudaMalloc(&ctx->stage_mask, T * HW));
    }
    float* dd = ctx->stage_depth + (size_t)slot * T * HW;
    char* df = reinterpret_cast<char*>(ctx->stage_flow) + (size_t)slot * T * ctx->flow_elems * fb;
    // the slot was last read kFrameRing steps ago on ctx->stream; copies run on the copy stream after that work
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done, 0));
    // one copy per plane when the caller's tracks are contiguous (the usual case), else one per track
    const bool depth_dense = (size_t)f->depth_track_stride == HW;
    const bool flow_dense = (size_t)f->flow_track_stride == ctx->flow_elems;
    bool all_masks = f->mask != nullptr && (size_t)f->mask_track_stride == HW;
    for (size_t t = 0; all_masks && f->mask_valid && t < T; ++t) all_masks = f->mask_valid[t] != 0;
    if (depth_dense) CK(cudaMemcpyAsync(dd, f->depth, sizeof(float) * T * HW, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (f->flow && flow_dense) CK(cudaMemcpyAsync(df, f->flow, T * ctx->flow_elems * fb, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (all_masks) CK(cudaMemcpyAsync(ctx->stage_mask, f->mask, T * HW, cudaMemcpyHostToDevice, ctx->copy_stream));
    for (size_t t = 0; t < T; ++t) {
        if (!depth_dense)
            CK(cudaMemcpyAsync(dd + t * HW, f->depth + t * f->depth_track_stride, sizeof(float) * HW, cudaMemcpyHostToDevice,
                               ctx->copy_stream));
        if (f->flow && !flow_dense)
            CK(cudaMemcpyAsync(df + t * ctx->flow_elems * fb,
                               reinterpret_cast<const char*>(f->flow) + t * f->flow_track_stride * fb, ctx->flow_elems * fb,
                               cudaMemcpyHostToDevice, ctx->copy_stream));
        if (f->mask && !all_masks && (!f->mask_valid || f->mask_valid[t]))
            CK(cudaMemcpyAsync(ctx->stage_mask + t * HW, f->mask + t * f->mask_track_stride, HW, cudaMemcpyHostToDevice,
                               ctx->copy_stream));
    }
    CK(cudaEventRecord(ctx->stage_event, ctx->copy_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->stage_event, 0));
    CK(cudaStreamWaitEvent(ctx->prep_stream, ctx->stage_event, 0));  // (the new mask is read on the prep / mask streams)
    // the caller may reuse its host buffers as soon as this call returns
    CK(cudaEventSynchronize(ctx->stage_event));
    *d_depth = dd;
    *d_flow = f->flow ? df : nullptr;
    *d_mask = f->mask ? ctx->stage_mask : nullptr;
    return 0;
}

// A hard (< 0) error after the host state machines have advanced leaves them ahead of the device work: the context is
// marked and every later step fails until roftb_filter_init() re-synchronises both sides.
static int filter_step_impl(roftb_ctx* ctx, const roftb_frame* f);

int roftb_filter_step(roftb_ctx* ctx, const roftb_frame* f) {
    if (!ctx || !f) return -2;
    if (ctx->comp) {  // the parts step side by side: the same frame with every per-track pointer shifted
        Composite& c = *ctx->comp;
        roftb_frame fr[kMaxParts];
        const size_t fb = ctx->g.flow_s16 ? 2 : 4;
        for (int h = 0; h < c.n; ++h) {
            const size_t o = (size_t)c.first[h];
            fr[h] = *f;
            if (f->depth) fr[h].depth = f->depth + o * (size_t)f->depth_track_stride;
            if (f->flow) fr[h].flow = reinterpret_cast<const char*>(f->flow) + o * (size_t)f->flow_track_stride * fb;
            if (f->mask) fr[h].mask = f->mask + o * (size_t)f->mask_track_stride;
            if (f->flow_valid) fr[h].flow_valid = f->flow_valid + o;
            if (f->mask_valid) fr[h].mask_valid = f->mask_valid + o;
            if (f->pose) fr[h].pose = f->pose + o * 7;
            if (f->pose_valid) fr[h].pose_valid = f->pose_valid + o;
            if (f->dt) fr[h].dt = f->dt + o;
        }
        const int rc = c.run_all([&](int h) { return roftb_filter_step(c.sub[h], &fr[h]); });
        return comp_fail(ctx, rc);
    }
    if (ctx->poisoned) return fail(ctx, "roftb_filter_step: a previous step failed; call roftb_filter_init first");
    const long long idx0 = ctx->frame_idx;
    const int rc = filter_step_impl(ctx, f);
    if (rc < 0 && ctx->frame_idx == idx0 && ctx->err.find("roftb_filter_step:") != 0 && ctx->err.find("device planes") != 0 &&
        ctx->err.find("track strides") != 0)
        ctx->poisoned = true;  // (argument errors are reported before anything is touched)
    return rc;
}

static int filter_step_impl(roftb_ctx* ctx, const roftb_frame* f) {
    if (!f->depth) return fail(ctx, "roftb_filter_step: depth is required (ROFTFilter.cpp:261-266 tears down without it)");
    const int T = ctx->T;
    const roftb_config& cfg = ctx->cfg;
    CK(cudaSetDevice(ctx->dev));
    const int slot = (int)(ctx->frame_idx % kFrameRing);
    const int hist_slot = (int)(ctx->frame_idx % kHistRing);
    const int cslot = (int)(ctx->frame_idx % kCtlRing);

    const float* d_depth = nullptr;
    const void* d_flow = nullptr;
    const uint8_t* d_mask = nullptr;
    long long depth_stride, flow_stride, mask_stride;
    if (f->memory == ROFTB_MEM_HOST) {
        int rc = stage_host_frame(ctx, f, slot, &d_depth, &d_flow, &d_mask);
        if (rc) return rc;
        depth_stride = (long long)ctx->HW;
        flow_stride = (long long)ctx->flow_elems;
        mask_stride = (long long)ctx->HW;
    } else {
        d_depth = f->depth; d_flow = f->flow; d_mask = f->mask;
        depth_stride = f->depth_track_stride; flow_stride = f->flow_track_stride; mask_stride = f->mask_track_stride;
        if ((((uintptr_t)d_depth) & 15) || (depth_stride & 3) || (d_flow && ((((uintptr_t)d_flow) & 15) || (flow_stride & 7))) ||
            (d_mask && ((((uintptr_t)d_mask) & 15) || (mask_stride & 15))))
            return fail(ctx, "device planes must be 16-byte aligned with 16-byte aligned track strides");
    }
    if (ctx->frame_idx > 0 && (ctx->ft.depth_stride != depth_stride || (d_flow && ctx->ft.flow_stride && ctx->ft.flow_stride != flow_stride)))
        return fail(ctx, "track strides must not change between frames");
    ctx->ft.depth[slot] = d_depth;
    ctx->ft.flow[slot] = d_flow;
    ctx->ft.depth_stride = depth_stride;
    if (d_flow) ctx->ft.flow_stride = flow_stride;

    // ---- host state machines -> control blocks ------------------------------------------------
    if (ctx->ctl_event_used[cslot]) CK(cudaEventSynchronize(ctx->ctl_event[cslot]));
    if (ctx->ukf_event_used[cslot]) CK(cudaEventSynchronize(ctx->ukf_event[cslot]));
    WarpCtl* wc = ctx->h_wctl + (size_t)cslot * T;
    VelCtl* vc = ctx->h_vctl + (size_t)cslot * T;
    UkfOp* ops = ctx->h_ops + (size_t)cslot * T * kMaxUkfOps;
    int32_t* nops = ctx->h_nops + (size_t)cslot * T;
    bool any_new_mask = false, any_vel = false, any_or = false, any_stage_before = false, any_stage_after = false;
    if (cfg.outlier_rejection && !ctx->mesh_nv) return fail(ctx, "outlier_rejection is enabled but no mesh was set (roftb_set_mesh)");
    for (int t = 0; t < T; ++t) {
        TrackHost& h = ctx->th[t];
        const bool flow_in = d_flow && (!f->flow_valid || f->flow_valid[t]);
        const bool mask_in = d_mask && (!f->mask_valid || f->mask_valid[t]);
        const bool pose_in = f->pose && (!f->pose_valid || f->pose_valid[t]);
        const double dt = f->dt ? f->dt[t] : cfg.sample_time;

        // -- ImageSegmentationOFAidedSource::step_frame (hpp:128-231), content-independent part
        WarpCtl& w = wc[t];
        memset(&w, 0, sizeof(w));
        w.flow_aided = cfg.flow_aided;
        w.cur_slot = slot;
        w.has_new = mask_in;
        if (cfg.flow_aided) {
            w.first_mask = mask_in && !h.seg_src_available;
            if (mask_in) h.seg_src_available = true;
            w.flow_valid = flow_in && !h.of_first_frame;
            h.of_first_frame = false;
            // ImageSegmentationMeasurement::freeze (cpp:56-68): new_segmentation_ = segmentation_available_
            if (h.seg_src_available) h.segmeas_available = true;
        } else {
            if (mask_in) h.segmeas_available = true;
        }
        any_new_mask |= mask_in;

        // -- ImageOpticalFlowMeasurement::freeze early-outs (hpp:184-229)
        VelCtl& v = vc[t];
        v.enable = 0;
        v.prev_slot = h.prev_slot < 0 ? slot : h.prev_slot;
        v.cur_slot = slot;
        v.hist_slot = hist_slot;
        v.dt = dt;
        if (h.segmeas_available) {
            if (!flow_in || h.fm_first_frame) {
                h.fm_first_frame = false;
            } else {
                v.enable = 1;
            }
            h.prev_slot = slot;  // previous_depth_ = depth (hpp:222,286)
        }
        any_vel |= v.enable != 0;

        // -- pose UKF sequencing (ROFTFilter.cpp:325-367) and CartesianQuaternionMeasurement::freeze
        PoseMeasHost& pm = h.pm;
        UkfOp* o = ops + (size_t)t * kMaxUkfOps;
        int n = 0;
        auto add_predict = [&]() {
            memset(&o[n], 0, sizeof(UkfOp));
            o[n].kind = kOpPredict;
            o[n].dt = dt;
            ++n;
        };
        auto add_correct = [&](int mtype, int vel_slot) {
            memset(&o[n], 0, sizeof(UkfOp));
            o[n].kind = kOpCorrect;
            o[n].meas_type = mtype;
            o[n].vel_slot = vel_slot;
            for (int i = 0; i < 7; ++i) o[n].meas[6 + i] = pm.last_pose[i];
            // a 13-sized measurement goes through correct_outlier_rejection (ROFTFilter.cpp:346-347, 357-358)
            if (cfg.outlier_rejection && mtype == ROFTB_MEAS_POSE_VELOCITY) {
                o[n].kind = kOpCorrectBoth;
                any_or = true;
            }
            ++n;
        };
        // Standard freeze (.cpp:176-347)
        if (cfg.use_velocity) {
            pm.is_first_velocity_in = true;
            pm.last_vel_slot = hist_slot;
        }
        pm.is_pose = false;
        if (cfg.use_pose && pose_in) {
            pm.is_pose = true;
            memcpy(pm.last_pose, f->pose + (size_t)t * 7, sizeof(double) * 7);
        }
        bool valid_freeze = true;
        if (pm.is_first_velocity_in && pm.is_pose) {
            pm.mtype = ROFTB_MEAS_POSE_VELOCITY;
            pm.meas_vel_slot = pm.last_vel_slot;
            pm.buffer.push_back(pm.last_vel_slot);
        } else if (pm.is_first_velocity_in) {
            pm.mtype = ROFTB_MEAS_VELOCITY;
            pm.meas_vel_slot = pm.last_vel_slot;
            pm.buffer.push_back(pm.last_vel_slot);
        } else if (pm.is_pose) {
            pm.mtype = ROFTB_MEAS_POSE;
        } else {
            pm.mtype = ROFTB_MEAS_NONE;
            valid_freeze = false;
        }
        // entries older than pose_delay + 1 are trimmed before any use (.cpp:100-104): cap the host mirror
        while ((int)pm.buffer.size() > ROFTB_MAX_DELAY + 3) pm.buffer.pop_front();
        if (!valid_freeze) {
            add_predict();  // p_corr_belief_ = p_pred_belief_
        } else if (pm.mtype == ROFTB_MEAS_POSE_VELOCITY && cfg.use_pose_resync) {
            // ROFTFilter.cpp:331-354: rewind to the buffered belief and replay the buffered velocities
            memset(&o[n], 0, sizeof(UkfOp));
            o[n].kind = kOpSwapBuffered;
            ++n;
            for (;;) {
                // PopBufferedMeasurement (.cpp:97-152)
                if (cfg.pose_delay > 0)
                    while ((int)pm.buffer.size() > cfg.pose_delay + 1) pm.buffer.pop_front();
                if (pm.buffer.empty()) {
                    pm.buffer.push_back(pm.meas_vel_slot);
                    break;
                }
                const int vs = pm.buffer.front();
                pm.buffer.pop_front();
                pm.last_vel_slot = vs;
                pm.meas_vel_slot = vs;
                if (pm.is_pose) {
                    pm.mtype = ROFTB_MEAS_POSE_VELOCITY;
                    pm.is_pose = false;
                } else {
                    pm.mtype = ROFTB_MEAS_VELOCITY;
                }
                if (n + 2 > kMaxUkfOps) return fail(ctx, "UKF op list overflow");
                add_predict();
                add_correct(pm.mtype, vs);
            }
        } else {
            add_predict();
            add_correct(pm.mtype, pm.mtype == ROFTB_MEAS_POSE ? -1 : pm.last_vel_slot);
        }
        nops[t] = n;
        // buffered features of the render-and-compare test -> UkfOp::pad of the track's first op: bit 2 = refresh before
        // this step's test (ROFTFilter.cpp:313-321: first step; :357-358: no re-sync, the test uses the current features),
        // bit 1 = refresh after it (:352-353: end of a re-synchronisation)
        if (cfg.outlier_rejection) {
            int bits = 0;
            const bool resync_now = o[0].kind == kOpSwapBuffered;
            bool tests = false;
            for (int i = 0; i < n; ++i) tests |= o[i].kind == kOpCorrectBoth;
            if (cfg.use_pose_resync) {
                if (!h.or_features_initialized) {
                    bits |= 2;
                    h.or_features_initialized = true;
                }
                if (resync_now) bits |= 1;
            } else if (tests) {
                bits |= 2;
            }
            o[0].pad = bits;
            any_stage_before |= (bits & 2) != 0;
            any_stage_after |= (bits & 1) != 0;
        }
    }

    cudaStream_t s = ctx->stream;
    const int par = (int)(ctx->frame_idx & 1);
    WarpCtl* d_wctl = ctx->d_wctl + (size_t)par * T;
    CK(cudaMemcpyAsync(d_wctl, wc, sizeof(WarpCtl) * T, cudaMemcpyHostToDevice, ctx->prep_stream));
    CK(cudaMemcpyAsync(ctx->d_vctl, vc, sizeof(VelCtl) * T, cudaMemcpyHostToDevice, s));
    UkfOp* d_ops = ctx->d_ops + (size_t)cslot * T * kMaxUkfOps;
    int32_t* d_nops = ctx->d_nops + (size_t)cslot * T;
    CK(cudaMemcpyAsync(d_ops, ops, sizeof(UkfOp) * T * kMaxUkfOps, cudaMemcpyHostToDevice, ctx->ukf_stream));
    CK(cudaMemcpyAsync(d_nops, nops, sizeof(int32_t) * T, cudaMemcpyHostToDevice, ctx->ukf_stream));
    if (cfg.outlier_rejection) {
        CK(cudaEventRecord(ctx->orj.ops_ready, ctx->ukf_stream));
        // a mask plane staged two steps ago is recycled by this step: the staging copy must be through
        const int old = (int)((ctx->frame_idx + 1) % 3);  // == (frame_idx - 2) mod 3
        if (ctx->frame_idx >= 2 && ctx->orj.stage_done_used[old]) {
            for (cudaStream_t st : {s, ctx->prep_stream, ctx->mask_stream}) CK(cudaStreamWaitEvent(st, ctx->orj.stage_done[old], 0));
            ctx->orj.stage_done_used[old] = false;
        }
    }
    CK(cudaEventRecord(ctx->ctl_event[cslot], s));
    // the pose UKF trails on its own stream; keep it within kUkfLag steps of the streaming kernels (velocity-history
    // ring): a pose re-sync replays pose_delay + 1 predict/correct pairs in one step, which the slack spreads out
    {
        const int lag = (cslot + kCtlRing - kUkfLag) % kCtlRing;
        if (ctx->ukf_event_used[lag]) CK(cudaStreamWaitEvent(s, ctx->ukf_event[lag], 0));
    }
    ctx->ctl_event_used[cslot] = true;

    // ---- device work ------------------------------------------------------------------------------
    // prep stream : mask stats / plan (needs nothing from the other streams), then - behind the previous step's
    //               velocity kernel and scatter - the initialisation of the destination plane of the tracks whose mask
    //               is NOT propagated by the velocity kernel, and the worklist of a newly delivered mask
    // main stream : the velocity kernel: worklist, pass A (+ propagation of the mask of the tracks that received no
    //               new one), pairing, median select, pass B, 6x6 epilogue - one cluster per track, one launch
    // mask stream : scatter/gather of the tracks with a NEW mask (chained through the buffered flows)
    // ukf stream  : pose UKF (needs only the twist published by the velocity kernel and the host-built op list)
    const uint8_t* seg_prev = ctx->mask_state[ctx->mask_cur];
    uint8_t* seg_next = ctx->mask_state[(ctx->mask_cur + 1) % 3];
    const uint8_t* occ_prev = ctx->mask_occ[ctx->mask_cur];
    uint8_t* occ_next = ctx->mask_occ[(ctx->mask_cur + 1) % 3];
    int32_t* wt_list = ctx->wt_list + (size_t)par * T * ctx->n_units;
    int32_t* wt_n = ctx->wt_n + (size_t)par * 2 * T;
    int32_t* nl_count = ctx->nl_count + (size_t)par * T * ctx->n_units;
    int32_t* nl_list = ctx->nl_list + (size_t)par * T * ctx->n_units;
    int32_t* nl_n = ctx->nl_n + (size_t)par * 2 * T;
    WarpPlan* plan = ctx->plan + (size_t)par * T;
    cudaEvent_t* pe = ctx->prof_on ? ctx->prof_ev[cslot] : nullptr;
    if (pe) prof_collect(ctx, cslot);
    MaskSyncArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.g = ctx->g; ma.ft = ctx->ft; ma.n_tracks = T;
    ma.new_mask = any_new_mask ? d_mask : nullptr; ma.new_stride = mask_stride;
    ma.state_src = seg_prev; ma.state_dst = seg_next; ma.winner = ctx->winner;
    ma.occ_src = occ_prev; ma.occ_dst = occ_next;
    ma.ctl = d_wctl; ma.stat = ctx->stat; ma.plan = plan; ma.fbuf = ctx->fbuf;
    ma.segm_delay = cfg.segm_delay;
    ma.s_list = wt_list; ma.s_n = wt_n; ma.n_list = nl_list; ma.n_n = nl_n; ma.n_warp_tiles = ctx->n_units;
    ma.fuse = 1;
    unsigned long long* span = nullptr;
    if (pe) {
        span = ctx->span_clock + (size_t)(ctx->frame_idx % 8) * 10;
        CK(cudaMemsetAsync(span, 0xFF, 10 * sizeof(unsigned long long), ctx->prep_stream));
    }
    ma.span_clock = span;
    {
        cudaStream_t ps = ctx->prep_stream;
        // worklist of a newly delivered mask; the same read gathers the statistics the plan needs
        if (any_new_mask && launch_tile_list(d_mask, mask_stride, 0, ctx->g.HW, T, nl_count, nl_list, nl_n,
                                             reinterpret_cast<const int32_t*>(d_wctl), (int)(sizeof(WarpCtl) / 4), ps, ctx->stat))
            return fail(ctx, "launch_tile_list failed");
        if (launch_mask_plan(ma, ps, true)) return fail(ctx, "launch_mask_plan failed");
        CK(cudaEventRecord(ctx->prep_event[par], ps));  // the velocity kernel only needs the plan
        // the planes this step initialises were last read by the previous step; a copied state was written by it
        if (ctx->vel_done_event_used) CK(cudaStreamWaitEvent(ps, ctx->vel_done_event, 0));
        if (ctx->mask_event_used) CK(cudaStreamWaitEvent(ps, ctx->mask_event, 0));
        if (pe) CK(cudaEventRecord(pe[10], ps));
        if (launch_mask_init(ma, ps)) return fail(ctx, "launch_mask_init failed");
        if (pe) CK(cudaEventRecord(pe[0], ps));
        CK(cudaEventRecord(ctx->prep2_event[par], ps));
    }
    CK(cudaStreamWaitEvent(s, ctx->prep_event[par], 0));
    if (ctx->mask_event_used) CK(cudaStreamWaitEvent(s, ctx->mask_event, 0));  // the previous step's scatter wrote seg_prev
    {
        cudaStream_t ms = ctx->mask_stream;
        CK(cudaStreamWaitEvent(ms, ctx->prep2_event[par], 0));
        if (pe) CK(cudaEventRecord(pe[6], ms));
        // worklist of the state mask for the (rare) tracks that propagate a mixed-valued mask outside the velocity kernel
        if (launch_flag_list(occ_prev, ctx->n_units, T, wt_list, wt_n, plan, ms)) return fail(ctx, "launch_flag_list failed");
        if (launch_mask_scatter_gather(ma, ms)) return fail(ctx, "launch_mask_scatter_gather failed");
        if (pe) CK(cudaEventRecord(pe[7], ms));
        CK(cudaEventRecord(ctx->mask_event, ms));
        ctx->mask_event_used = true;
        ctx->mask_cur = (ctx->mask_cur + 1) % 3;
    }
    {
        VelocityArgs a;
        memset(&a, 0, sizeof(a));
        a.g = ctx->g; a.ft = ctx->ft; a.n_tracks = T;
        a.seg = seg_prev; a.seg_stride = (long long)ctx->HW; a.thr = 1;  // cv::threshold(> 1) applied on load
        a.occ_src = occ_prev;
        a.ctl = ctx->d_vctl; a.weight_flow = cfg.weight_flow;
        a.scratch = ctx->scratch;
        a.v_mean = ctx->v_mean; a.v_cov = ctx->v_cov; a.q_diag = ctx->q_diag;
        a.r_flow[0] = cfg.cov_flow[0]; a.r_flow[1] = cfg.cov_flow[1];
        a.fx = cfg.fx; a.fy = cfg.fy; a.cx = cfg.cx; a.cy = cfg.cy;
        a.accum_fp64 = cfg.accum_fp64;
        a.vel_hist = ctx->vel_hist; a.hist_ring = kHistRing;
        a.out_count = ctx->d_count; a.out_lambda = ctx->d_lambda; a.out_eta = ctx->d_eta;
        a.wl_units = ctx->wl_units; a.wl_pixels = ctx->wl_pixels;
        a.phase_clock = pe ? ctx->phase_clock : nullptr;
        a.span_clock = span ? span + 6 : nullptr;
        a.order = ctx->vel_order + (size_t)par * T; a.order_next = ctx->vel_order + (size_t)(par ^ 1) * T;
        a.done_ticket = ctx->vel_ticket;
        a.side_stream[0] = ctx->vel_side[0]; a.side_stream[1] = ctx->vel_side[1];
        a.side_fork = ctx->vel_fork; a.side_join[0] = ctx->vel_join[0]; a.side_join[1] = ctx->vel_join[1];
        a.update_state = 1;
        a.fuse_scatter = 1; a.plan = plan; a.state_dst = seg_next; a.occ_dst = occ_next;
        if (pe) CK(cudaEventRecord(pe[1], s));
        if (launch_velocity(a, s)) return fail(ctx, "launch_velocity failed");
        if (pe) CK(cudaEventRecord(pe[2], s));
        CK(cudaEventRecord(ctx->vel_done_event, s));
        ctx->vel_done_event_used = true;
        CK(cudaEventRecord(ctx->vel_event[cslot], s));
        if (any_new_mask) {
            // the launch order the velocity kernel left for the next step describes the masks BEFORE this delivery:
            // redo it from the flags of the new state once both writers of that state are done
            cudaStream_t ms = ctx->mask_stream;
            CK(cudaStreamWaitEvent(ms, ctx->vel_done_event, 0));
            if (launch_order_from_flags(occ_next, ctx->n_units, T, ctx->order_units, ctx->vel_ticket + 1, a.order_next, ms)) return fail(ctx, "launch_order_from_flags failed");
            CK(cudaEventRecord(ctx->mask_event, ms));
        }
    }
    const bool any_stage = any_stage_before || any_stage_after;
    if (any_stage) {
        // stage <- (mask state after this step, depth of this step) of the flagged tracks, once both writers of the state
        // (velocity kernel, new-mask scatter) are done; the previous staging must have been consumed
        auto& oj = ctx->orj;
        cudaStream_t cs = oj.stream;
        CK(cudaStreamWaitEvent(cs, oj.ops_ready, 0));
        CK(cudaStreamWaitEvent(cs, ctx->vel_done_event, 0));
        CK(cudaStreamWaitEvent(cs, ctx->mask_event, 0));
        if (oj.unstage_done_used) CK(cudaStreamWaitEvent(cs, oj.unstage_done, 0));
        if (launch_or_features(ctx->g, T, d_ops, kMaxUkfOps, 3, seg_next, (long long)ctx->HW, 1, d_depth, depth_stride, oj.wt_count,
                               oj.rank_total, oj.feat_stage, oj.feat_stride, oj.n_stage, cs))
            return fail(ctx, "launch_or_features failed");
        const int cur = (int)(ctx->frame_idx % 3);
        CK(cudaEventRecord(oj.stage_done[cur], cs));
        oj.stage_done_used[cur] = true;
    }
    {
        cudaStream_t us = ctx->ukf_stream;
        CK(cudaStreamWaitEvent(us, ctx->vel_event[cslot], 0));
        UkfArgs a;
        memset(&a, 0, sizeof(a));
        a.n_tracks = T; a.p = ctx->ukf_p;
        a.ops = d_ops; a.n_ops = d_nops; a.max_ops = kMaxUkfOps;
        a.mean = ctx->p_mean; a.cov = ctx->p_cov; a.buf_mean = ctx->pb_mean; a.buf_cov = ctx->pb_cov;
        a.vel_hist = ctx->vel_hist; a.hist_ring = kHistRing;
        a.span_clock = span ? span + 8 : nullptr;
        static const bool warm_on = [] { const char* e = getenv("ROFTB_UKF_WARM"); return !e || atoi(e) != 0; }();
        a.warm = warm_on ? ctx->ukf_warm : nullptr;
        auto& oj = ctx->orj;
        if (cfg.outlier_rejection) {
            a.resume = oj.resume; a.cand_mean = oj.cand_mean; a.cand_cov = oj.cand_cov;
            CK(cudaMemsetAsync(oj.resume, 0, sizeof(int32_t) * T, us));
        }
        if (pe) CK(cudaEventRecord(pe[8], us));
        if (launch_ukf(a, us)) return fail(ctx, "launch_ukf failed");  // everything, or up to the first render-and-compare test
        const int stage_slot = (int)(ctx->frame_idx % 3);
        if (any_stage_before) {
            CK(cudaStreamWaitEvent(us, oj.stage_done[stage_slot], 0));
            if (launch_or_feat_copy(T, d_ops, kMaxUkfOps, 2, oj.feat_stage, oj.n_stage, oj.feat_snap, oj.n_snap, oj.feat_stride, us))
                return fail(ctx, "launch_or_feat_copy failed");
        }
        if (any_or) {
            // ROFTFilter::pick_best_alternative (ROFTFilter.cpp:467-621) on the buffered features, all on this stream
            const int dv = oj.divider;
            const size_t tile = (size_t)(ctx->g.W / dv) * (ctx->g.H / dv);
            if (launch_or_models(T, oj.resume, oj.cand_mean, oj.model, us)) return fail(ctx, "launch_or_models failed");
            RenderArgs ra;
            ra.n_items = 2 * T;
            ra.vertices = ctx->mesh_vertices; ra.n_vertices = ctx->mesh_nv;
            ra.faces = ctx->mesh_faces; ra.n_faces = ctx->mesh_nf;
            ra.model = oj.model;
            ra.fx = (float)(cfg.fx / dv); ra.fy = (float)(cfg.fy / dv); ra.cx = (float)(cfg.cx / dv); ra.cy = (float)(cfg.cy / dv);
            ra.w = ctx->g.W / dv; ra.h = ctx->g.H / dv;
            ra.scale = ctx->mesh_scale; ra.n_scale = T;
            if (launch_render_depth(ra, oj.vertex_scratch, oj.zbuf, oj.rendered, us)) return fail(ctx, "launch_render_depth failed");
            if (launch_or_l1(T, oj.resume, oj.feat_snap, oj.n_snap, oj.feat_stride, oj.rendered, (long long)tile, dv, ctx->g.W, oj.err,
                             oj.samples, us))
                return fail(ctx, "launch_or_l1 failed");
            if (launch_pick_best(T, oj.err, oj.samples, cfg.outlier_rejection_gain, oj.selected, nullptr, us)) return fail(ctx, "launch_pick_best failed");
            if (launch_or_select(T, oj.resume, oj.selected, oj.cand_mean, oj.cand_cov, ctx->p_mean, ctx->p_cov, us))
                return fail(ctx, "launch_or_select failed");
            if (launch_ukf(a, us)) return fail(ctx, "launch_ukf failed");  // the rest of the op lists
        }
        if (any_stage_after) {
            CK(cudaStreamWaitEvent(us, oj.stage_done[stage_slot], 0));
            if (launch_or_feat_copy(T, d_ops, kMaxUkfOps, 1, oj.feat_stage, oj.n_stage, oj.feat_snap, oj.n_snap, oj.feat_stride, us))
                return fail(ctx, "launch_or_feat_copy failed");
        }
        if (any_stage) {
            CK(cudaEventRecord(oj.unstage_done, us));
            oj.unstage_done_used = true;
        }
        if (pe) {
            CK(cudaEventRecord(pe[9], us));
            ctx->prof_used[cslot] = true;
        }
        CK(cudaEventRecord(ctx->ukf_event[cslot], us));
        ctx->ukf_event_used[cslot] = true;
    }
    // staged host frames may be overwritten once BOTH the main and the mask stream are done with them
    if (f->memory == ROFTB_MEM_HOST) {
        CK(cudaStreamWaitEvent(s, ctx->mask_event, 0));
        CK(cudaEventRecord(ctx->copy_done, s));
    }
    (void)any_vel;
    ctx->frame_idx++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return -1; }
    return 0;
}

int roftb_get_state(roftb_ctx* ctx, double* p_mean, double* p_cov, double* v_mean, double* v_cov) {
    if (!ctx) return -2;
    if (ctx->comp) {
        Composite& c = *ctx->comp;
        int rc = 0;
        for (int h = 0; h < c.n; ++h) {
            const size_t o = (size_t)c.first[h];
            rc |= roftb_get_state(c.sub[h], p_mean ? p_mean + o * 13 : nullptr, p_cov ? p_cov + o * 144 : nullptr, v_mean ? v_mean + o * 6 : nullptr,
                                  v_cov ? v_cov + o * 36 : nullptr);
        }
        return comp_fail(ctx, rc);
    }
    const size_t T = ctx->T;
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    CK(cudaEventRecord(ctx->join_event, ctx->ukf_stream));
    CK(cudaStreamWaitEvent(s, ctx->join_event, 0));
    if (p_mean) CK(cudaMemcpyAsync(p_mean, ctx->p_mean, T * 13 * 8, cudaMemcpyDeviceToHost, s));
    if (p_cov) CK(cudaMemcpyAsync(p_cov, ctx->p_cov, T * 144 * 8, cudaMemcpyDeviceToHost, s));
    if (v_mean) CK(cudaMemcpyAsync(v_mean, ctx->v_mean, T * 6 * 8, cudaMemcpyDeviceToHost, s));
    if (v_cov) CK(cudaMemcpyAsync(v_cov, ctx->v_cov, T * 36 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

int roftb_get_mask(roftb_ctx* ctx, uint8_t* raw, uint8_t* thresholded) {
    if (!ctx) return -2;
    if (ctx->comp) {
        Composite& c = *ctx->comp;
        int rc = 0;
        for (int h = 0; h < c.n; ++h) {
            const size_t o = (size_t)c.first[h] * ctx->HW;
            rc |= roftb_get_mask(c.sub[h], raw ? raw + o : nullptr, thresholded ? thresholded + o : nullptr);
        }
        return comp_fail(ctx, rc);
    }
    const size_t n = (size_t)ctx->T * ctx->HW;
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    const uint8_t* cur = ctx->mask_state[ctx->mask_cur];
    if (ctx->mask_event_used) CK(cudaStreamWaitEvent(s, ctx->mask_event, 0));
    if (raw) CK(cudaMemcpyAsync(raw, cur, n, cudaMemcpyDeviceToHost, s));
    if (thresholded) {
        if (!ctx->thr_tmp) CK(cudaMalloc(&ctx->thr_tmp, n));
        if (launch_threshold(cur, ctx->thr_tmp, n, s)) return fail(ctx, "launch_threshold failed");
        CK(cudaMemcpyAsync(thresholded, ctx->thr_tmp, n, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return 0;
}

int roftb_get_velocity_info(roftb_ctx* ctx, int32_t* count, double* lambda, double* eta) {
    if (!ctx) return -2;
    if (ctx->comp) {
        Composite& c = *ctx->comp;
        int rc = 0;
        for (int h = 0; h < c.n; ++h) {
            const size_t o = (size_t)c.first[h];
            rc |= roftb_get_velocity_info(c.sub[h], count ? count + o : nullptr, lambda ? lambda + o * 36 : nullptr, eta ? eta + o * 6 : nullptr);
        }
        return comp_fail(ctx, rc);
    }
    const size_t T = ctx->T;
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    if (count) CK(cudaMemcpyAsync(count, ctx->d_count, T * 4, cudaMemcpyDeviceToHost, s));
    if (lambda) CK(cudaMemcpyAsync(lambda, ctx->d_lambda, T * 36 * 8, cudaMemcpyDeviceToHost, s));
    if (eta) CK(cudaMemcpyAsync(eta, ctx->d_eta, T * 6 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

int roftb_get_worklist(roftb_ctx* ctx, int32_t* units, int32_t* pixels) {
    if (!ctx) return -2;
    if (ctx->comp) {
        Composite& c = *ctx->comp;
        int rc = 0;
        for (int h = 0; h < c.n; ++h) rc |= roftb_get_worklist(c.sub[h], units ? units + c.first[h] : nullptr, pixels ? pixels + c.first[h] : nullptr);
        return comp_fail(ctx, rc);
    }
    if (ctx->frame_idx == 0) return fail(ctx, "roftb_get_worklist: no step yet");
    const size_t T = ctx->T;
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    if (units) CK(cudaMemcpyAsync(units, ctx->wl_units, T * 4, cudaMemcpyDeviceToHost, s));
    if (pixels) CK(cudaMemcpyAsync(pixels, ctx->wl_pixels, T * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

}  // extern "C"

// =============================================================================================
// Stateless operators: host buffers in, host buffers out, through temporary device memory.
// =============================================================================================
namespace {

struct TmpBuf {
    std::vector<void*> ptrs;
    ~TmpBuf() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <class T>
    T* alloc(size_t n, bool zero = false) {
        void* p = nullptr;
        if (cudaMalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
        if (zero) cudaMemset(p, 0, (n ? n : 1) * sizeof(T));
        ptrs.push_back(p);
        return reinterpret_cast<T*>(p);
    }
    template <class T>
    T* upload(const T* h, size_t n, cudaStream_t s) {
        T* d = alloc<T>(n);
        if (d && h) cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, s);
        return d;
    }
};

}  // namespace

extern "C" {

int roftb_mask_sync(roftb_ctx* ctx, int32_t n_masks, const uint8_t* mask, const void* flows, int32_t n_flows,
                    int32_t zero_origin, uint8_t* out_raw, uint8_t* out_thr) {
    if (ctx && ctx->comp) return comp_fail(ctx, roftb_mask_sync(ctx->comp->sub[0], n_masks, mask, flows, n_flows, zero_origin, out_raw, out_thr));  // operators run on the first half context
    if (!ctx || !mask || n_masks <= 0 || n_flows < 0 || n_flows > kMaxChain || (n_flows > 0 && !flows))
        return ctx ? fail(ctx, "roftb_mask_sync: bad argument") : -2;
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    const size_t HW = ctx->HW, N = n_masks;
    const size_t fb = flow_scalar_bytes(ctx);
    TmpBuf tb;
    uint8_t* d_mask = tb.upload(mask, N * HW, s);
    char* d_flows = n_flows ? tb.upload(reinterpret_cast<const char*>(flows), (size_t)n_flows * N * ctx->flow_elems * fb, s) : nullptr;
    uint8_t* d_out = tb.alloc<uint8_t>(N * HW);
    uint8_t* d_thr = tb.alloc<uint8_t>(N * HW);
    int32_t* d_win = tb.alloc<int32_t>(N * HW);
    std::vector<WarpPlan> plans(N);
    for (size_t i = 0; i < N; ++i) {
        // hpp:211-226 with the buffer given explicitly
        const uint8_t* m = mask + i * HW;
        WarpPlan& p = plans[i];
        memset(&p, 0, sizeof(p));
        p.mode = kWarpScatter;
        p.src_new = 1;
        p.zero_origin = zero_origin ? 1 : 0;
        p.n_flows = n_flows;
        for (int j = 0; j < n_flows; ++j) p.flow_slot[j] = j;
        int vmin = 256, vmax = 0;
        for (size_t k = (zero_origin ? 1 : 0); k < HW; ++k)
            if (m[k]) { vmin = m[k] < vmin ? m[k] : vmin; vmax = m[k] > vmax ? m[k] : vmax; }
        p.uniform_val = (vmax > 0 && vmin == vmax) ? vmin : 0;
        p.dflt = zero_origin ? 0 : m[0];
    }
    WarpPlan* d_plan = tb.upload(plans.data(), N, s);
    if (!d_mask || !d_out || !d_thr || !d_win || !d_plan || (n_flows && !d_flows)) return fail(ctx, "roftb_mask_sync: out of device memory");
    MaskSyncArgs a;
    memset(&a, 0, sizeof(a));
    a.g = ctx->g; a.n_tracks = (int)N;
    for (int j = 0; j < n_flows; ++j) a.ft.flow[j] = d_flows + (size_t)j * N * ctx->flow_elems * fb;
    a.ft.flow_stride = (long long)ctx->flow_elems;
    a.new_mask = d_mask; a.new_stride = (long long)HW;
    a.state_src = d_mask; a.state_dst = d_out; a.winner = d_win; a.plan = d_plan;
    a.occ_src = nullptr; a.occ_dst = nullptr;
    {
        const int nwt = ctx->n_units;
        int32_t* cnt = tb.alloc<int32_t>(N * nwt);
        int32_t* lst = tb.alloc<int32_t>(N * nwt);
        int32_t* ln = tb.alloc<int32_t>(2 * N);
        if (!cnt || !lst || !ln) return fail(ctx, "roftb_mask_sync: out of device memory");
        if (launch_tile_list(d_mask, (long long)HW, 0, ctx->g.HW, (int)N, cnt, lst, ln, nullptr, 0, s)) return fail(ctx, "launch_tile_list failed");
        a.s_list = lst; a.s_n = ln; a.n_list = lst; a.n_n = ln; a.n_warp_tiles = nwt;
    }
    if (launch_mask_sync(a, s, true)) return fail(ctx, "launch_mask_sync failed");
    if (out_raw) CK(cudaMemcpyAsync(out_raw, d_out, N * HW, cudaMemcpyDeviceToHost, s));
    if (out_thr) {
        if (launch_threshold(d_out, d_thr, N * HW, s)) return fail(ctx, "launch_threshold failed");
        CK(cudaMemcpyAsync(out_thr, d_thr, N * HW, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return 0;
}

static int velocity_operator(roftb_ctx* ctx, int32_t n, const uint8_t* mask, const float* depth, const void* flow,
                             const double* x_pred, const double* dt, double* x, double* P, double* lambda, double* eta,
                             int32_t* count, bool update) {
    if (!ctx || n <= 0 || !mask || !depth || !flow) return ctx ? fail(ctx, "velocity operator: bad argument") : -2;
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    const size_t HW = ctx->HW, N = n;
    const size_t fb = flow_scalar_bytes(ctx);
    TmpBuf tb;
    uint8_t* d_mask = tb.upload(mask, N * HW, s);
    float* d_depth = tb.upload(depth, N * HW, s);
    char* d_flow = tb.upload(reinterpret_cast<const char*>(flow), N * ctx->flow_elems * fb, s);
    std::vector<VelCtl> ctl(N);
    for (size_t i = 0; i < N; ++i) {
        ctl[i].enable = 1; ctl[i].prev_slot = 0; ctl[i].cur_slot = 0; ctl[i].hist_slot = -1;
        ctl[i].dt = dt ? dt[i] : ctx->cfg.sample_time;
    }
    VelCtl* d_ctl = tb.upload(ctl.data(), N, s);
    std::vector<double> zeros(N * 36, 0.0);
    double* d_x = tb.upload(x ? x : zeros.data(), N * 6, s);
    double* d_P = tb.upload(P ? P : zeros.data(), N * 36, s);
    double* d_xp = x_pred ? tb.upload(x_pred, N * 6, s) : nullptr;
    const int nwt = ctx->n_units;
    VelocityArgs a;
    memset(&a, 0, sizeof(a));
    a.g = ctx->g; a.n_tracks = (int)N;
    a.ft.depth[0] = d_depth; a.ft.flow[0] = d_flow;
    a.ft.depth_stride = (long long)HW; a.ft.flow_stride = (long long)ctx->flow_elems;
    a.seg = d_mask; a.seg_stride = (long long)HW; a.thr = 0;  // previous_segmentation_ is used through findNonZero
    uint8_t* d_occ = tb.alloc<uint8_t>(N * nwt);
    a.occ_src = d_occ;
    a.ctl = d_ctl; a.weight_flow = ctx->cfg.weight_flow;
    // the scratch pool of the context (this operator runs on the context's main stream); per-item arrays are sized here
    a.scratch = ctx->scratch;
    a.scratch.track_slot = tb.alloc<int32_t>(N);
    a.scratch.chunk_aux = tb.alloc<int32_t>(N * kVtMaxChunks);
    a.scratch.part = tb.alloc<double>(N * kVtMaxCluster * kVtPartN);
    a.scratch.track_sel = tb.alloc<double>(N * 4);
    a.v_mean = d_x; a.v_cov = d_P; a.q_diag = ctx->q_diag;
    a.r_flow[0] = ctx->cfg.cov_flow[0]; a.r_flow[1] = ctx->cfg.cov_flow[1];
    a.fx = ctx->cfg.fx; a.fy = ctx->cfg.fy; a.cx = ctx->cfg.cx; a.cy = ctx->cfg.cy;
    a.accum_fp64 = ctx->cfg.accum_fp64;
    a.out_count = tb.alloc<int32_t>(N);
    a.out_lambda = tb.alloc<double>(N * 36);
    a.out_eta = tb.alloc<double>(N * 6);
    a.update_state = update ? 1 : 0;
    a.x_pred_override = d_xp;
    if (!d_mask || !d_depth || !d_flow || !d_ctl || !d_x || !d_P || !d_occ || !a.scratch.track_slot || !a.scratch.chunk_aux ||
        !a.scratch.part || !a.scratch.track_sel ||
        !a.out_count || !a.out_lambda || !a.out_eta)
        return fail(ctx, "velocity operator: out of device memory");
    if (launch_unit_flags(d_mask, (long long)HW, ctx->g.HW, (int)N, d_occ, s)) return fail(ctx, "launch_unit_flags failed");
    if (launch_velocity(a, s)) return fail(ctx, "launch_velocity failed");
    if (update && x) CK(cudaMemcpyAsync(x, d_x, N * 6 * 8, cudaMemcpyDeviceToHost, s));
    if (update && P) CK(cudaMemcpyAsync(P, d_P, N * 36 * 8, cudaMemcpyDeviceToHost, s));
    if (lambda) CK(cudaMemcpyAsync(lambda, a.out_lambda, N * 36 * 8, cudaMemcpyDeviceToHost, s));
    if (eta) CK(cudaMemcpyAsync(eta, a.out_eta, N * 6 * 8, cudaMemcpyDeviceToHost, s));
    if (count) CK(cudaMemcpyAsync(count, a.out_count, N * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

int roftb_flow_velocity(roftb_ctx* ctx, int32_t n_items, const uint8_t* mask, const float* depth, const void* flow,
                        const double* x_pred, const double* dt, double* lambda, double* eta, int32_t* count) {
    if (ctx && ctx->comp) return comp_fail(ctx, roftb_flow_velocity(ctx->comp->sub[0], n_items, mask, depth, flow, x_pred, dt, lambda, eta, count));  // operators run on the first half context
    return velocity_operator(ctx, n_items, mask, depth, flow, x_pred, dt, nullptr, nullptr, lambda, eta, count, false);
}

int roftb_velocity_kf(roftb_ctx* ctx, int32_t n_items, const uint8_t* mask, const float* depth, const void* flow,
                      const double* dt, double* x, double* P, int32_t* count) {
    if (ctx && ctx->comp) return comp_fail(ctx, roftb_velocity_kf(ctx->comp->sub[0], n_items, mask, depth, flow, dt, x, P, count));  // operators run on the first half context
    if (!x || !P) return ctx ? fail(ctx, "roftb_velocity_kf: x and P are required") : -2;
    return velocity_operator(ctx, n_items, mask, depth, flow, nullptr, dt, x, P, nullptr, nullptr, count, true);
}

static int ukf_operator(roftb_ctx* ctx, int32_t n, double* mean, double* cov, const std::vector<UkfOp>& ops) {
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    const size_t N = n;
    TmpBuf tb;
    double* d_mean = tb.upload(mean, N * 13, s);
    double* d_cov = tb.upload(cov, N * 144, s);
    UkfOp* d_ops = tb.upload(ops.data(), N, s);
    std::vector<int32_t> nops(N, 1);
    int32_t* d_nops = tb.upload(nops.data(), N, s);
    if (!d_mean || !d_cov || !d_ops || !d_nops) return fail(ctx, "ukf operator: out of device memory");
    UkfArgs a;
    memset(&a, 0, sizeof(a));
    a.n_tracks = n; a.p = ctx->ukf_p; a.ops = d_ops; a.n_ops = d_nops; a.max_ops = 1; a.mean = d_mean; a.cov = d_cov;
    if (launch_ukf(a, s)) return fail(ctx, "launch_ukf failed");
    CK(cudaMemcpyAsync(mean, d_mean, N * 13 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(cov, d_cov, N * 144 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

int roftb_ukf_predict(roftb_ctx* ctx, int32_t n_items, double* mean, double* cov, const double* dt) {
    if (ctx && ctx->comp) return comp_fail(ctx, roftb_ukf_predict(ctx->comp->sub[0], n_items, mean, cov, dt));  // operators run on the first half context
    if (!ctx || n_items <= 0 || !mean || !cov) return ctx ? fail(ctx, "roftb_ukf_predict: bad argument") : -2;
    std::vector<UkfOp> ops(n_items);
    for (int i = 0; i < n_items; ++i) {
        memset(&ops[i], 0, sizeof(UkfOp));
        ops[i].kind = kOpPredict;
        ops[i].dt = dt ? dt[i] : ctx->cfg.sample_time;
    }
    return ukf_operator(ctx, n_items, mean, cov, ops);
}

int roftb_ukf_correct(roftb_ctx* ctx, int32_t n_items, double* mean, double* cov, const double* meas, const int32_t* meas_type) {
    if (ctx && ctx->comp) return comp_fail(ctx, roftb_ukf_correct(ctx->comp->sub[0], n_items, mean, cov, meas, meas_type));  // operators run on the first half context
    if (!ctx || n_items <= 0 || !mean || !cov || !meas || !meas_type) return ctx ? fail(ctx, "roftb_ukf_correct: bad argument") : -2;
    std::vector<UkfOp> ops(n_items);
    for (int i = 0; i < n_items; ++i) {
        memset(&ops[i], 0, sizeof(UkfOp));
        ops[i].kind = kOpCorrect;
        ops[i].meas_type = meas_type[i];
        ops[i].vel_slot = -1;
        memcpy(ops[i].meas, meas + (size_t)i * 13, sizeof(double) * 13);
    }
    return ukf_operator(ctx, n_items, mean, cov, ops);
}

static void fill_select(roftb_ctx* ctx, SelectArgs& a, TmpBuf& tb, int32_t n, const uint8_t* mask, const float* depth,
                        const void* flow, cudaStream_t s) {
    const size_t HW = ctx->HW, N = n;
    const size_t fb = flow_scalar_bytes(ctx);
    memset(&a, 0, sizeof(a));
    a.g = ctx->g; a.n_items = n;
    a.mask = tb.upload(mask, N * HW, s); a.mask_stride = (long long)HW; a.thr = 0;
    a.depth = tb.upload(depth, N * HW, s); a.depth_stride = (long long)HW;
    a.flow = flow ? tb.upload(reinterpret_cast<const char*>(flow), N * ctx->flow_elems * fb, s) : nullptr;
    a.flow_stride = (long long)ctx->flow_elems;
    a.wt_count = tb.alloc<int32_t>(N * ctx->n_warp_tiles);
    a.wt_count2 = tb.alloc<int32_t>(N * ctx->n_warp_tiles);
}

int roftb_flow_measurement_export(roftb_ctx* ctx, const uint8_t* mask, const float* depth, const void* flow, double dt,
                                  int32_t capacity, double* z, double* H, int32_t* n_valid) {
    if (ctx && ctx->comp) return comp_fail(ctx, roftb_flow_measurement_export(ctx->comp->sub[0], mask, depth, flow, dt, capacity, z, H, n_valid));  // operators run on the first half context
    if (!ctx || !mask || !depth || !flow || capacity < 0 || !n_valid) return ctx ? fail(ctx, "export: bad argument") : -2;
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    TmpBuf tb;
    SelectArgs a;
    fill_select(ctx, a, tb, 1, mask, depth, flow, s);
    double* d_z = tb.alloc<double>((size_t)capacity * 2);
    double* d_H = tb.alloc<double>((size_t)capacity * 12);
    int32_t* d_n = tb.alloc<int32_t>(1, true);
    if (!a.mask || !a.depth || !a.flow || !a.wt_count || !a.wt_count2 || !d_z || !d_H || !d_n) return fail(ctx, "export: out of device memory");
    if (launch_export_measurement(a, dt, ctx->cfg.fx, ctx->cfg.fy, ctx->cfg.cx, ctx->cfg.cy, capacity, d_z, d_H, d_n, s)) return fail(ctx, "export launch failed");
    CK(cudaMemcpyAsync(n_valid, d_n, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int nv = *n_valid < capacity ? *n_valid : capacity;
    if (z && nv) CK(cudaMemcpy(z, d_z, (size_t)nv * 2 * 8, cudaMemcpyDeviceToHost));
    if (H && nv) CK(cudaMemcpy(H, d_H, (size_t)nv * 12 * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int roftb_masked_points(roftb_ctx* ctx, int32_t n_items, const uint8_t* mask, const float* depth, double max_depth,
                        int32_t capacity, double* points, int32_t* count) {
    if (ctx && ctx->comp) return comp_fail(ctx, roftb_masked_points(ctx->comp->sub[0], n_items, mask, depth, max_depth, capacity, points, count));  // operators run on the first half context
    if (!ctx || n_items <= 0 || !mask || !depth || capacity < 0 || !count) return ctx ? fail(ctx, "masked_points: bad argument") : -2;
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    TmpBuf tb;
    SelectArgs a;
    fill_select(ctx, a, tb, n_items, mask, depth, nullptr, s);
    double* d_pts = tb.alloc<double>((size_t)n_items * capacity * 3);
    int32_t* d_n = tb.alloc<int32_t>(n_items, true);
    if (!a.mask || !a.depth || !a.wt_count || !a.wt_count2 || !d_pts || !d_n) return fail(ctx, "masked_points: out of device memory");
    if (launch_masked_points(a, max_depth, ctx->cfg.fx, ctx->cfg.fy, ctx->cfg.cx, ctx->cfg.cy, capacity, d_pts, d_n, s))
        return fail(ctx, "masked_points launch failed");
    CK(cudaMemcpyAsync(count, d_n, (size_t)n_items * 4, cudaMemcpyDeviceToHost, s));
    if (points && capacity) CK(cudaMemcpyAsync(points, d_pts, (size_t)n_items * capacity * 3 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

int roftb_masked_depth_l1(roftb_ctx* ctx, int32_t n_items, const uint8_t* mask, const float* depth, const float* rendered,
                          int32_t divider, double* err_sum, int32_t* samples) {
    if (ctx && ctx->comp) return comp_fail(ctx, roftb_masked_depth_l1(ctx->comp->sub[0], n_items, mask, depth, rendered, divider, err_sum, samples));  // operators run on the first half context
    if (!ctx || n_items <= 0 || !mask || !depth || !rendered || divider <= 0 || !err_sum || !samples)
        return ctx ? fail(ctx, "masked_depth_l1: bad argument") : -2;
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    TmpBuf tb;
    SelectArgs a;
    fill_select(ctx, a, tb, n_items, mask, depth, nullptr, s);
    const size_t rsz = (size_t)(ctx->g.W / divider) * (ctx->g.H / divider);
    float* d_r = tb.upload(rendered, (size_t)n_items * rsz, s);
    double* d_e = tb.alloc<double>(n_items, true);
    int32_t* d_n = tb.alloc<int32_t>(n_items, true);
    if (!a.mask || !a.depth || !a.wt_count || !d_r || !d_e || !d_n) return fail(ctx, "masked_depth_l1: out of device memory");
    if (launch_masked_depth_l1(a, d_r, (long long)rsz, divider, d_e, d_n, s)) return fail(ctx, "masked_depth_l1 launch failed");
    CK(cudaMemcpyAsync(err_sum, d_e, (size_t)n_items * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(samples, d_n, (size_t)n_items * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

}  // extern "C"

// ---- pose outlier rejection ------------------------------------------------------------------
namespace {

// Model matrix of SICAD::superimpose (SICAD.cpp:604-607) for a pose (x, y, z, axis, angle): glm::rotate(I, angle, axis)
// in float (UPSTREAM-RECALL of glm: normalised axis, Rodrigues with c = cos(angle), s = sin(angle)), translation in the
// last column.  out: rotation row-major (9), translation (3).
void model_matrix(const double* pose7, float* out) {
    const float ang = static_cast<float>(pose7[6]);
    float ax = static_cast<float>(pose7[3]), ay = static_cast<float>(pose7[4]), az = static_cast<float>(pose7[5]);
    const float c = std::cos(ang), s = std::sin(ang);
    const float n = std::sqrt(ax * ax + ay * ay + az * az);
    if (n > 0.f) { ax /= n; ay /= n; az /= n; }
    const float tx = (1.f - c) * ax, ty = (1.f - c) * ay, tz = (1.f - c) * az;
    out[0] = c + tx * ax;      out[1] = ty * ax - s * az;  out[2] = tz * ax + s * ay;
    out[3] = tx * ay + s * az; out[4] = c + ty * ay;       out[5] = tz * ay - s * ax;
    out[6] = tx * az - s * ay; out[7] = ty * az + s * ax;  out[8] = c + tz * az;
    out[9] = static_cast<float>(pose7[0]); out[10] = static_cast<float>(pose7[1]); out[11] = static_cast<float>(pose7[2]);
}

// Eigen::AngleAxisd(Eigen::Quaterniond(w, x, y, z)) as used at ROFTFilter.cpp:518-524 (UPSTREAM-RECALL of Eigen 3.3+:
// angle = 2 atan2(|v|, |w|), axis = v / (+-|v|) with the sign of w; identity -> angle 0, axis (1, 0, 0))
void state_to_pose7(const double* state13, double* pose7) {
    pose7[0] = state13[6]; pose7[1] = state13[7]; pose7[2] = state13[8];
    const double w = state13[9], x = state13[10], y = state13[11], z = state13[12];
    double n = std::sqrt(x * x + y * y + z * z);
    if (n != 0.0) {
        pose7[6] = 2.0 * std::atan2(n, std::fabs(w));
        if (w < 0.0) n = -n;
        pose7[3] = x / n; pose7[4] = y / n; pose7[5] = z / n;
    } else {
        pose7[6] = 0.0; pose7[3] = 1.0; pose7[4] = 0.0; pose7[5] = 0.0;
    }
}

int render_tiles(roftb_ctx* ctx, TmpBuf& tb, int n_items, const double* poses7, int divider, float** d_out, cudaStream_t s) {
    const int w = ctx->g.W / divider, h = ctx->g.H / divider;
    std::vector<float> model((size_t)n_items * 12);
    for (int i = 0; i < n_items; ++i) model_matrix(poses7 + (size_t)i * 7, &model[(size_t)i * 12]);
    RenderArgs ra;
    ra.n_items = n_items;
    ra.vertices = ctx->mesh_vertices; ra.n_vertices = ctx->mesh_nv;
    ra.faces = ctx->mesh_faces; ra.n_faces = ctx->mesh_nf;
    ra.model = tb.upload(model.data(), model.size(), s);
    // SICAD is constructed with every intrinsic divided by divider_ (ROFTFilter.cpp:194-197)
    ra.fx = (float)(ctx->cfg.fx / divider); ra.fy = (float)(ctx->cfg.fy / divider);
    ra.cx = (float)(ctx->cfg.cx / divider); ra.cy = (float)(ctx->cfg.cy / divider);
    ra.w = w; ra.h = h;
    ra.scale = nullptr; ra.n_scale = 1;  // (the operators render the mesh as uploaded)
    void* vs = tb.alloc<char>(render_vertex_scratch_bytes(n_items, ctx->mesh_nv));
    uint32_t* zb = tb.alloc<uint32_t>((size_t)n_items * w * h);
    float* out = tb.alloc<float>((size_t)n_items * w * h);
    if (!ra.model || !vs || !zb || !out) return -1;
    cudaStreamSynchronize(s);  // `model` is a local vector
    if (launch_render_depth(ra, vs, zb, out, s)) return -1;
    *d_out = out;
    return 0;
}

}  // namespace

extern "C" {

int roftb_set_mesh(roftb_ctx* ctx, const float* vertices, int32_t n_vertices, const int32_t* faces, int32_t n_faces) {
    if (ctx && ctx->comp) {
        int rc = 0;
        for (int h = 0; h < ctx->comp->n; ++h) rc |= roftb_set_mesh(ctx->comp->sub[h], vertices, n_vertices, faces, n_faces);
        return comp_fail(ctx, rc);
    }
    if (!ctx || !vertices || !faces || n_vertices <= 0 || n_faces <= 0) return ctx ? fail(ctx, "roftb_set_mesh: bad argument") : -2;
    for (int64_t i = 0; i < (int64_t)n_faces * 3; ++i)
        if (faces[i] < 0 || faces[i] >= n_vertices) return fail(ctx, "roftb_set_mesh: face index out of range");
    CK(cudaSetDevice(ctx->dev));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->mesh_vertices) cudaFree(ctx->mesh_vertices);
    if (ctx->mesh_faces) cudaFree(ctx->mesh_faces);
    ctx->mesh_vertices = nullptr; ctx->mesh_faces = nullptr; ctx->mesh_nv = ctx->mesh_nf = 0;
    CK(cudaMalloc(&ctx->mesh_vertices, (size_t)n_vertices * 3 * sizeof(float)));
    CK(cudaMalloc(&ctx->mesh_faces, (size_t)n_faces * 3 * sizeof(int32_t)));
    CK(cudaMemcpy(ctx->mesh_vertices, vertices, (size_t)n_vertices * 3 * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->mesh_faces, faces, (size_t)n_faces * 3 * sizeof(int32_t), cudaMemcpyHostToDevice));
    ctx->mesh_nv = n_vertices; ctx->mesh_nf = n_faces;
    if (ctx->cfg.outlier_rejection) {
        CK(cudaStreamSynchronize(ctx->ukf_stream));
        if (ctx->orj.vertex_scratch) cudaFree(ctx->orj.vertex_scratch);
        ctx->orj.vertex_scratch = nullptr;
        CK(cudaMalloc(&ctx->orj.vertex_scratch, render_vertex_scratch_bytes(2 * ctx->T, n_vertices)));
    }
    return 0;
}

int roftb_set_mesh_scale(roftb_ctx* ctx, const float* scale) {
    if (!ctx) return -2;
    if (ctx->comp) {
        int rc = 0;
        for (int h = 0; h < ctx->comp->n; ++h) rc |= roftb_set_mesh_scale(ctx->comp->sub[h], scale ? scale + (size_t)ctx->comp->first[h] * 3 : nullptr);
        return comp_fail(ctx, rc);
    }
    CK(cudaSetDevice(ctx->dev));
    CK(cudaStreamSynchronize(ctx->ukf_stream));
    if (!scale) {
        if (ctx->mesh_scale) cudaFree(ctx->mesh_scale);
        ctx->mesh_scale = nullptr;
        return 0;
    }
    if (!ctx->mesh_scale) CK(cudaMalloc(&ctx->mesh_scale, (size_t)ctx->T * 3 * sizeof(float)));
    CK(cudaMemcpy(ctx->mesh_scale, scale, (size_t)ctx->T * 3 * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

int roftb_render_depth(roftb_ctx* ctx, int32_t n_items, const double* poses7, int32_t divider, float* out_depth) {
    if (ctx && ctx->comp) return comp_fail(ctx, roftb_render_depth(ctx->comp->sub[0], n_items, poses7, divider, out_depth));  // operators run on the first half context
    if (!ctx || n_items <= 0 || !poses7 || divider <= 0 || !out_depth) return ctx ? fail(ctx, "roftb_render_depth: bad argument") : -2;
    if (!ctx->mesh_nv) return fail(ctx, "roftb_render_depth: no mesh (roftb_set_mesh)");
    if (ctx->g.W % divider || ctx->g.H % divider) return fail(ctx, "roftb_render_depth: divider must divide the frame size");
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    TmpBuf tb;
    float* d_out = nullptr;
    if (render_tiles(ctx, tb, n_items, poses7, divider, &d_out, s)) return fail(ctx, "roftb_render_depth: launch failed / out of device memory");
    const size_t tile = (size_t)(ctx->g.W / divider) * (ctx->g.H / divider);
    CK(cudaMemcpyAsync(out_depth, d_out, (size_t)n_items * tile * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

int roftb_pick_best_alternative(roftb_ctx* ctx, int32_t n_items, const uint8_t* segmentation, const float* depth,
                                const double* alternatives, int32_t divider, double gain, int32_t* selected, double* likelihoods) {
    if (ctx && ctx->comp) return comp_fail(ctx, roftb_pick_best_alternative(ctx->comp->sub[0], n_items, segmentation, depth, alternatives, divider, gain, selected, likelihoods));  // operators run on the first half context
    if (!ctx || n_items <= 0 || !segmentation || !depth || !alternatives || divider <= 0 || !selected || !(gain != 0.0))
        return ctx ? fail(ctx, "roftb_pick_best_alternative: bad argument") : -2;
    if (!ctx->mesh_nv) return fail(ctx, "roftb_pick_best_alternative: no mesh (roftb_set_mesh)");
    if (ctx->g.W % divider || ctx->g.H % divider) return fail(ctx, "roftb_pick_best_alternative: divider must divide the frame size");
    CK(cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    TmpBuf tb;
    // tiles laid out [alternative][item]: one masked-L1 launch per alternative over the same masks and depths
    std::vector<double> poses((size_t)2 * n_items * 7);
    for (int k = 0; k < 2; ++k)
        for (int i = 0; i < n_items; ++i) state_to_pose7(alternatives + ((size_t)i * 2 + k) * 13, &poses[((size_t)k * n_items + i) * 7]);
    float* d_r = nullptr;
    if (render_tiles(ctx, tb, 2 * n_items, poses.data(), divider, &d_r, s)) return fail(ctx, "roftb_pick_best_alternative: render failed");
    SelectArgs a;
    fill_select(ctx, a, tb, n_items, segmentation, depth, nullptr, s);
    const size_t rsz = (size_t)(ctx->g.W / divider) * (ctx->g.H / divider);
    double* d_e = tb.alloc<double>((size_t)2 * n_items, true);
    int32_t* d_n = tb.alloc<int32_t>((size_t)2 * n_items, true);
    int32_t* d_sel = tb.alloc<int32_t>(n_items);
    double* d_l = tb.alloc<double>((size_t)2 * n_items);
    if (!a.mask || !a.depth || !a.wt_count || !d_e || !d_n || !d_sel || !d_l) return fail(ctx, "roftb_pick_best_alternative: out of device memory");
    for (int k = 0; k < 2; ++k)
        if (launch_masked_depth_l1(a, d_r + (size_t)k * n_items * rsz, (long long)rsz, divider, d_e + (size_t)k * n_items,
                                   d_n + (size_t)k * n_items, s))
            return fail(ctx, "roftb_pick_best_alternative: L1 launch failed");
    if (launch_pick_best(n_items, d_e, d_n, gain, d_sel, d_l, s)) return fail(ctx, "roftb_pick_best_alternative: launch failed");
    CK(cudaMemcpyAsync(selected, d_sel, (size_t)n_items * 4, cudaMemcpyDeviceToHost, s));
    if (likelihoods) CK(cudaMemcpyAsync(likelihoods, d_l, (size_t)2 * n_items * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

}  // extern "C"
