// Optical-flow-aided velocity measurement + linear Kalman correction, batched over tracks
// (north-star part 1).
//
// Replaces, for every track at once:
//   ImageOpticalFlowMeasurement<T>::freeze        src/roft-lib/include/ROFT/ImageOpticalFlowMeasurement.hpp:231-283
//   SKFCorrection::correctStep                    src/roft-lib/src/SKFCorrection.cpp:37-153
//   SpatialVelocityModel + bfl::KFPrediction      src/roft-lib/src/SpatialVelocityModel.cpp:15-27
//   observability gate                            src/roft-lib/src/ROFTFilter.cpp:294-301
//
// The reference materialises z (2N) and H (2N x 6) and then runs a SEQUENTIAL 2-row Kalman update per
// pixel.  For per-pixel independent noise that is algebraically the information-form sum
//     Lambda = (P+Q)^-1 + sum_j l_j H_j^T R^-1 H_j ,  eta = (P+Q)^-1 x + sum_j l_j H_j^T R^-1 z_j
// (SURVEY.md F1), so the per-pixel work becomes a streaming reduction.  With x^ = (u-cx)/fx,
// y^ = (v-cy)/fy, a = 1/d the two rows of H_j are  dt*fx*L1 and dt*fy*L2  with
//     L1 = [a, 0, -x^a, -x^y^, 1+x^2, -y^]     L2 = [0, a, -y^a, -(1+y^2), x^y^, x^]
// so the kernel accumulates S1 = sum l L1^T L1, S2 = sum l L2^T L2, g1 = sum l L1^T dx,
// g2 = sum l L2^T dy in FP32 per thread, tree-reduces with warp shuffles, and hands FP64 block
// partials to a per-track epilogue that applies the FP64 constants and solves the 6x6 system.
// No atomics in the accumulation path.
//
// Laplacian re-weighting (SKFCorrection.cpp:91-116) needs the median of the innovation norms: pass A
// streams the listed units of the frame once and writes the norms to position-addressed slots (128 per
// listed unit, -1 = not a valid measurement); an exact 3-level radix select (12 + 12 + 8 key bits) finds
// the middle order statistic(s), its last pass also gathers what b = mean|n - m| needs; pass B streams the
// listed units again (depth, flow and the pass-A norms) and accumulates with the weights.
// Two implementations of the passes: k_flow_pass_ring (TMA-staged, the common configuration) and the
// generic register-prefetch k_flow_pass (everything else); see DESIGN.md 4.
#include "roftb_internal.cuh"

namespace roftb {
namespace {

__device__ __forceinline__ float4 get4(const float4* p, bool ok) { return ok ? ld_nc_f4(p) : make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// ---- per-warp-tile counts of selected candidates (only needed when stride > 1 or for ordered output) ----
__global__ void __launch_bounds__(kThreads) k_mask_count(const uint8_t* __restrict__ seg, long long seg_stride, int thr,
                                                        int HW, int n_warp_tiles, int32_t* __restrict__ wt_count,
                                                        const VelCtl* __restrict__ ctl) {
    const int t = blockIdx.y;
    if (ctl && !ctl[t].enable) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t thr4 = (uint32_t)thr * 0x01010101u;
    const uint32_t* mq = reinterpret_cast<const uint32_t*>(seg + (long long)t * seg_stride);
    const int nq = HW >> 2;
    for (int wt = blockIdx.x * (kThreads / 32) + warp; wt < n_warp_tiles; wt += gridDim.x * (kThreads / 32)) {
        int c = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = wt * 128 + j * 32 + lane;
            const uint32_t m = q < nq ? ld_nc_u32(mq + q) : 0u;
            c += __popc(__vcmpgtu4(m, thr4)) >> 3;
        }
        c = warp_sum(c);
        if (lane == 0) wt_count[(long long)t * n_warp_tiles + wt] = c;
    }
}

// in-place exclusive scan of each track's warp-tile counts (one block per track)
__global__ void __launch_bounds__(kThreads) k_wt_scan(int32_t* __restrict__ wt_count, int n_warp_tiles, int32_t* __restrict__ total,
                                                     const VelCtl* __restrict__ ctl) {
    const int t = blockIdx.x;
    if (ctl && !ctl[t].enable) return;
    __shared__ int sh[kThreads / 32];
    __shared__ int carry;
    int32_t* p = wt_count + (long long)t * n_warp_tiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_warp_tiles; base += kThreads) {
        const int i = base + threadIdx.x;
        const int v = i < n_warp_tiles ? p[i] : 0;
        int incl = warp_scan_incl(v, lane);
        if (lane == 31) sh[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += sh[w];
        const int c = carry;
        if (i < n_warp_tiles) p[i] = c + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == kThreads - 1) carry = c + woff + incl;
        __syncthreads();
    }
    if (total && threadIdx.x == 0) total[t] = carry;
}

// ---- worklist of non-empty units ------------------------------------------------------------------
// Masks cover a compact fraction of the frame, so every streaming kernel iterates over the list of non-empty
// UNITS of its track (a unit = 128 consecutive pixels = one quad per lane of a warp) instead of the whole plane:
// no time is spent on empty pixels, lanes are (almost) all busy inside a unit, and the work is evenly spread over
// the blocks of a track regardless of where the object is.
// k_tile_count packs, per unit, (#bytes > 0) | (#bytes > thr) << 16.
__global__ void __launch_bounds__(kThreads) k_tile_count(const uint8_t* __restrict__ plane, long long stride, int thr, int HW,
                                                        int n_units, int32_t* __restrict__ wt_count,
                                                        const int32_t* __restrict__ active, int active_stride) {
    const int t = blockIdx.y;
    if (active && !active[(long long)t * active_stride]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t thr4 = (uint32_t)thr * 0x01010101u;
    const uint32_t* mq = reinterpret_cast<const uint32_t*>(plane + (long long)t * stride);
    const uint4* m4 = reinterpret_cast<const uint4*>(plane + (long long)t * stride);
    const int n16 = HW >> 4;
    const int n_blk = (n_units + 3) >> 2;  // 512-px blocks: one 128-bit load per lane, 8 lanes per unit
    const unsigned gmask = 0xffu << (8 * (lane >> 3));
    auto count16 = [&](const uint4& w) {
        const int c0 = __popc(__vcmpne4(w.x, 0u)) + __popc(__vcmpne4(w.y, 0u)) + __popc(__vcmpne4(w.z, 0u)) + __popc(__vcmpne4(w.w, 0u));
        const int c1 = __popc(__vcmpgtu4(w.x, thr4)) + __popc(__vcmpgtu4(w.y, thr4)) + __popc(__vcmpgtu4(w.z, thr4)) +
                       __popc(__vcmpgtu4(w.w, thr4));
        return (unsigned)((c0 >> 3) | ((c1 >> 3) << 16));
    };
    const int wstride = gridDim.x * (kThreads / 32);
    for (int b = blockIdx.x * (kThreads / 32) + warp; b < n_blk; b += 2 * wstride) {
        // two independent 128-bit loads in flight per lane
        const int b2 = b + wstride;
        const int qa = b * 32 + lane, qb = b2 * 32 + lane;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        const uint4 wa = qa < n16 ? ld_nc_u4(m4 + qa) : zero;
        const uint4 wb = (b2 < n_blk && qb < n16) ? ld_nc_u4(m4 + qb) : zero;
        const unsigned ca = __reduce_add_sync(gmask, count16(wa));  // both 16-bit fields stay <= 128
        const unsigned cb = __reduce_add_sync(gmask, count16(wb));
        if ((lane & 7) == 0) {
            const int ua = b * 4 + (lane >> 3), ub = b2 * 4 + (lane >> 3);
            if (ua < n_units) wt_count[(long long)t * n_units + ua] = (int)ca;
            if (b2 < n_blk && ub < n_units) wt_count[(long long)t * n_units + ub] = (int)cb;
        }
    }
}

// one block per track: wt_count <- exclusive prefix of the (> thr) counts (row-major rank base), wt_list <- ids of
// the units holding any non-zero byte (ascending), wt_n <- their number
constexpr int kCompactThreads = 1024;
__global__ void __launch_bounds__(kCompactThreads) k_tile_compact(int32_t* __restrict__ wt_count, int32_t* __restrict__ wt_list,
                                                          int32_t* __restrict__ wt_n, int n_units,
                                                          const int32_t* __restrict__ active, int active_stride) {
    const int t = blockIdx.x;
    if (active && !active[(long long)t * active_stride]) return;
    __shared__ int sh_r[kCompactThreads / 32], sh_l[kCompactThreads / 32];
    __shared__ int carry_r, carry_l;
    int32_t* cnt = wt_count + (long long)t * n_units;
    int32_t* list = wt_list + (long long)t * n_units;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_r = 0; carry_l = 0; }
    __syncthreads();
    for (int base = 0; base < n_units; base += kCompactThreads) {
        const int i = base + threadIdx.x;
        const int packed = i < n_units ? cnt[i] : 0;
        const int vr = packed >> 16;
        const int vl = (packed & 0xffff) ? 1 : 0;
        const int ir = warp_scan_incl(vr, lane), il = warp_scan_incl(vl, lane);
        if (lane == 31) { sh_r[warp] = ir; sh_l[warp] = il; }
        __syncthreads();
        int wr = 0, wl = 0;
        for (int w = 0; w < warp; ++w) { wr += sh_r[w]; wl += sh_l[w]; }
        const int cr = carry_r, cl = carry_l;
        if (i < n_units) {
            cnt[i] = cr + wr + ir - vr;
            if (vl) list[cl + wl + il - 1] = i;
        }
        __syncthreads();
        if (threadIdx.x == kCompactThreads - 1) { carry_r = cr + wr + ir; carry_l = cl + wl + il; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        wt_n[t] = carry_l;              // number of non-empty units
        wt_n[gridDim.x + t] = carry_r;  // number of candidates (bytes > thr) of the whole plane
    }
}

// ---- the streaming pass ---------------------------------------------------------------------
// PASS 0 (A): innovation norms -> compact list.   PASS 1 (B): weighted normal-equation accumulation.
// FAST: float2 flow at full resolution (vector loads); otherwise generic per-pixel flow fetch.
struct PassArgs {
    Geom g;
    FrameTable ft;
    const uint8_t* seg; long long seg_stride; int thr;
    const VelCtl* ctl;
    int n_units;
    const int32_t* wt_prefix;   // [T][n_units] row-major rank base of each unit
    const int32_t* wt_list;     // [T][n_units] non-empty units
    const int32_t* wt_n;        // [T]
    long long norm_stride;      // norm slots per track
    float* norms; uint32_t* norm_count;
    const WeightParams* wp; int weight_flow;
    const double* x_pred; int x_stride;
    double* partials; int max_blocks;
    double fx, fy;
    double cxd, cyd, inv_fxd, inv_fyd;   // FP64 intrinsics: pixel -> normalised coordinates without systematic rounding
    // fused single-flow mask propagation (ImageSegmentationOFAidedSource.hpp:221-226) for tracks whose plan says so
    const WarpPlan* plan; uint8_t* state_dst; int32_t* winner;
    // accumulation-precision routing of pass B: -1 = every track; otherwise this launch handles the tracks whose
    // candidate count is (auto_small_is64 ? below : at or above) auto_threshold
    int auto_threshold; int auto_take_small;
    int reuse_norms;  // pass B, stride 1, weighting on: read the innovation norm / validity written by pass A
};

// AT = accumulation type of pass B: float (per-pixel terms and partial sums in FP32) or double (per-pixel terms
// and sums in FP64: forward error ~ cond(Lambda) * 1e-16 instead of cond * 1e-7 / sqrt(N)).
//
// Each warp walks its track's worklist of non-empty units, four list entries (4 x 128 px, one quad per lane each) per
// trip.  The
// sub-tile loop is kept ROLLED with the loads of the next quad in flight while the current one is processed: the
// kernel stays a few hundred SASS instructions (a fully unrolled body was 8k instructions and stalled on
// instruction fetch), registers stay below 128 and three 128-bit loads per lane are always outstanding.
// Innovation norms are written to position-addressed slots (list position x 128 + pixel; one coalesced 128-bit store
// per lane, gated-out / non-candidate pixels marked with -1) at stride 1, and to rank-addressed compact slots (rank base
// of the unit from the worklist prefix) when stride > 1; either way pass A needs no atomics.
// SCATTER: the same walk also forward-scatters every non-zero mask pixel through the current flow (the "no new mask"
// propagation of the mask synchronisation) - it needs exactly the mask words and flow values this pass loads anyway.
//
// The per-pixel body is written for instruction count (the kernel is issue-bound, not DRAM-bound): predicates instead
// of branches, one multiply to pack the per-byte compare result into a nibble, unsigned range checks, lane-private
// norm slots (one warp scan per tile instead of a ballot per pixel), row/column tracked incrementally instead of a
// division per quad.
__device__ __forceinline__ uint32_t nibble_of(uint32_t bytemask) {
    // bytemask has 0xff / 0x00 per byte (vcmp result): gather bit 0 of each byte into bits 0..3
    return ((bytemask & 0x01010101u) * 0x10204080u) >> 28;
}

template <int PASS, bool FAST, typename AT, bool SCATTER>
__global__ void __launch_bounds__(kThreads, PASS == 0 ? 3 : 2) k_flow_pass(PassArgs a) {
    const int t = blockIdx.y;
    const VelCtl c = a.ctl[t];
    bool do_sc = false;
    uint8_t sc_val = 0;
    if (SCATTER) {
        const WarpPlan& p = a.plan[t];
        do_sc = p.fused != 0;  // only single-valued masks are fused (WarpPlan::fused)
        sc_val = (uint8_t)p.uniform_val;
    }
    if (!c.enable && !do_sc) return;
    if (PASS == 1 && a.auto_threshold >= 0) {
        const bool small = a.wt_n[gridDim.y + t] < a.auto_threshold;
        if (small != (a.auto_take_small != 0)) return;  // the other precision variant handles this track
    }
    const Geom& g = a.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t thr4 = (uint32_t)a.thr * 0x01010101u;
    const uint32_t* mq = reinterpret_cast<const uint32_t*>(a.seg + (long long)t * a.seg_stride);
    const float4* dq = reinterpret_cast<const float4*>(a.ft.depth[c.prev_slot] + (long long)t * a.ft.depth_stride);
    const char* fbase = reinterpret_cast<const char*>(a.ft.flow[c.cur_slot]) +
                        (long long)t * a.ft.flow_stride * (g.flow_s16 ? 2 : 4);
    const float4* fq = reinterpret_cast<const float4*>(fbase);
    const int nq = g.HW >> 2;
    const int W = g.W;
    const unsigned uW = (unsigned)g.W, uH = (unsigned)g.H;
    const unsigned ustride = (unsigned)g.stride;
    uint8_t* dst_t = SCATTER ? a.state_dst + (long long)t * g.HW : nullptr;
    float* norms_t = a.norms + (long long)t * a.norm_stride;
    const float inv_fx = g.inv_fx, max_d = g.max_depth_f;
    const float inv_w = 1.0f / (float)g.W, Wf = (float)g.W, Hf = (float)g.H;
    const bool small_hw = g.HW < (1 << 24);
    const unsigned sc_bias = 0x4b000000u * (uW + 1u);  // see the scatter index below

    // predicted velocity (F = I: the predicted mean is the previous corrected mean) and FP32 row scales
    float x[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = (float)a.x_pred[(long long)t * a.x_stride + i];
    const float c1 = (float)(a.fx * c.dt), c2 = (float)(a.fy * c.dt);
    WeightParams wp;
    wp.use = 0;
    if (PASS == 1 && a.weight_flow) wp = a.wp[t];

    AT acc[kNAcc];
    // FP32 variant: S1/S2 (and g1/g2) are accumulated as packed pairs with FFMA2 (two FP32 FMAs per issue slot on sm_100)
    constexpr bool kPacked = (PASS == 1 && sizeof(AT) == 4);
    float2 acc2[kPacked ? 20 : 1];
    float cnt_acc = 0.f;
    if (PASS == 1) {
#pragma unroll
        for (int i = 0; i < kNAcc; ++i) acc[i] = (AT)0;
#pragma unroll
        for (int i = 0; i < (kPacked ? 20 : 1); ++i) acc2[i] = make_float2(0.f, 0.f);
    }

    const int n_list = a.wt_n[t];
    const int32_t* list = a.wt_list + (long long)t * a.n_units;
    const int32_t* prefix = a.wt_prefix + (long long)t * a.n_units;
    // each warp iteration takes four consecutive list entries (units of 128 px, one quad per lane each).
    // The dependent chain list -> unit id -> mask word -> depth / flow / norm loads would cost three exposed memory
    // latencies per group (a quarter of all stall samples in the ncu capture): unit ids are fetched two groups ahead
    // by lanes 0..3, and every line the NEXT group will touch is pulled into L2 (prefetch.global.L2: no registers)
    // while the current group is processed, so the chain only ever sees L2-hit latency.
    const int gstep = gridDim.x * (kThreads / 32) * 4;
    const int g_first = (blockIdx.x * (kThreads / 32) + warp) * 4;
    const bool reuse = PASS == 1 && a.reuse_norms;
    const float4* nq4 = reinterpret_cast<const float4*>(norms_t);
    int my_unit = -1, my_rank = 0, nx_unit = -1, nx2_unit = -1;
    if (lane < 4) {
        if (g_first + lane < n_list) my_unit = list[g_first + lane];
        if (g_first + gstep + lane < n_list) nx_unit = list[g_first + gstep + lane];
    }
#pragma unroll 1
    for (int g0 = g_first; g0 < n_list; g0 += gstep) {
        // lanes 0..3 hold the unit ids (and rank bases); broadcast by shuffle inside the rolled loop
        if (lane < 4 && g0 + 2 * gstep + lane < n_list) nx2_unit = list[g0 + 2 * gstep + lane];
        if (g.stride > 1 && my_unit >= 0) my_rank = prefix[my_unit];
        // candidate bits of this lane's quad in each unit: bit (4j + i); scs: non-zero pixels to propagate
        uint32_t sel = 0, scs = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int unit = __shfl_sync(0xffffffffu, my_unit, j);
            const int q = unit * 32 + lane;
            const uint32_t m = (unit >= 0 && q < nq) ? ld_nc_u32(mq + q) : 0u;
            sel |= nibble_of(__vcmpgtu4(m, thr4)) << (4 * j);
            if (SCATTER) {
                uint32_t z = nibble_of(__vcmpne4(m, 0u));
                if (q == 0) z &= ~1u;  // mask_(0,0) = 0 (hpp:224)
                scs |= z << (4 * j);
            }
        }
        if (!c.enable) sel = 0;
        if (!do_sc) scs = 0;
        const uint32_t need = sel | scs;
        if (g0 + gstep < n_list) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int unit = __shfl_sync(0xffffffffu, nx_unit, j);
                const int q = unit * 32 + lane;
                if (unit >= 0 && q < nq) {
                    prefetch_l2(mq + q);
                    prefetch_l2(dq + q);
                    if (FAST) {
                        prefetch_l2(fq + 2 * q);
                        prefetch_l2(fq + 2 * q + 1);
                    }
                    if (reuse) prefetch_l2(nq4 + (g0 + gstep + j) * 32 + lane);
                }
            }
        }

        // software pipeline over the four units: loads of unit j+1 are issued before unit j is processed
        // (two register sets used alternately - the rolled loop below runs two units per trip - so no copies)
        float4 DA, F0A, F1A, DB, F0B, F1B;
        float4 NA = make_float4(-1.f, -1.f, -1.f, -1.f), NB = NA;  // pass B: norms of pass A (negative: not a valid measurement)
        {
            const int q = __shfl_sync(0xffffffffu, my_unit, 0) * 32 + lane;
            if (reuse) NA = get4(nq4 + g0 * 32 + lane, (sel & 0xfu) != 0u);
            DA = get4(dq + q, (sel & 0xfu) != 0u);
            if (FAST) {
                const bool on = (need & 0xfu) != 0u;
                F0A = get4(fq + 2 * q, on);
                F1A = get4(fq + 2 * q + 1, on);
            }
        }
        auto unit_step = [&](const int j, const float4& Dc, const float4& F0c, const float4& F1c, const float4& Nc, float4& Dn,
                             float4& F0n, float4& F1n, float4& Nn) -> bool {
            if (j < 3) {
                const int qn = __shfl_sync(0xffffffffu, my_unit, j + 1) * 32 + lane;
                if (reuse) Nn = get4(nq4 + (g0 + j + 1) * 32 + lane, ((sel >> (4 * (j + 1))) & 0xfu) != 0u);
                Dn = get4(dq + qn, ((sel >> (4 * (j + 1))) & 0xfu) != 0u);
                if (FAST) {
                    const bool on = ((need >> (4 * (j + 1))) & 0xfu) != 0u;
                    F0n = get4(fq + 2 * qn, on);
                    F1n = get4(fq + 2 * qn + 1, on);
                }
            }
            const int unit = __shfl_sync(0xffffffffu, my_unit, j);
            if (unit < 0) return false;  // warp-uniform: past the end of the list
            uint32_t nib = (sel >> (4 * j)) & 0xfu;
            const uint32_t snib = (scs >> (4 * j)) & 0xfu;
            int nslot = 0;  // stride > 1: compact norm slot of this lane's first selected candidate
            if (g.stride > 1) {
                // keep rank % stride == 0 over the row-major rank of the candidates (hpp:237), BEFORE the gates
                const int cnt = __popc(nib);
                const int incl = warp_scan_incl(cnt, lane);
                unsigned rank = (unsigned)(__shfl_sync(0xffffffffu, my_rank, j) + incl - cnt);
                nslot = (int)((rank + ustride - 1u) / ustride);
                uint32_t ns = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if ((nib >> i) & 1u) {
                        if (rank % ustride == 0u) ns |= 1u << i;
                        ++rank;
                    }
                }
                nib = ns;
            }
            float4 nv = make_float4(-1.f, -1.f, -1.f, -1.f);  // pass A: norms of this quad (-1: not a valid measurement)
            if ((nib | snib) != 0u) {
                const int px = (unit * 32 + lane) << 2;
                // row / column of the quad: reciprocal multiply with a +-1 fix-up (exact for HW < 2^24; a 32-bit integer
                // division is ~35 instructions per quad)
                int v, u0;
                if (small_hw) {
                    v = (int)((float)px * inv_w);
                    u0 = px - v * W;
                    if (u0 < 0) { u0 += W; --v; }
                    if (u0 >= W) { u0 -= W; ++v; }
                } else {
                    v = px / W;
                    u0 = px - v * W;
                }
                const float vf = (float)v, u0f = (float)u0;
                float yh, xh0;
                double yhd = 0.0, xh0d = 0.0;
                if (PASS == 1) {
                    // normalised coordinates from FP64 intrinsics (a rounded 1/fx would bias every pixel the same way)
                    yhd = ((double)v - a.cyd) * a.inv_fyd;
                    xh0d = ((double)u0 - a.cxd) * a.inv_fxd;
                    yh = (float)yhd;
                    xh0 = (float)xh0d;
                } else {
                    // pass A only feeds the Laplacian weights: FP32 coordinates are plenty
                    yh = (vf - g.cy) * g.inv_fy;
                    xh0 = (u0f - g.cx) * inv_fx;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool cand = (nib >> i) & 1u;
                    float dx, dy;
                    if (FAST) {
                        const float4 f = i < 2 ? F0c : F1c;
                        dx = (i & 1) ? f.z : f.x;  // FAST: float2 flow, grid 1, scale 1
                        dy = (i & 1) ? f.w : f.y;
                    } else {
                        dx = 0.f;
                        dy = 0.f;
                        if (cand || (SCATTER && ((snib >> i) & 1u))) {
                            const float2 f = load_flow(fbase, (long long)(v / g.grid) * g.Wf + ((u0 + i) / g.grid), g);
                            dx = f.x;
                            dy = f.y;
                        }
                    }
                    if (SCATTER) {
                        // hpp:249-278 for a single flow: the source pixel is inside the frame and its flow element is the
                        // one just fetched; IEEE adds; C truncation (a NaN would convert to 0 on the GPU, so it is tested;
                        // +-inf / huge values saturate outside the frame exactly like x86's INT_MIN)
                        // In-frame test on the floats: (int)t in [0, n) <=> -1 < t < n (truncation; NaN fails both).
                        // trunc(max(t, 0)) without the conversion pipe: t + 2^23 rounded toward zero keeps floor(t) in the
                        // mantissa, so the float bits are 0x4b000000 + floor(t); the bias of both terms is folded into one
                        // constant (arithmetic mod 2^32).
                        const float tx = __fadd_rn(u0f + (float)i, dx), ty = __fadd_rn(vf, dy);
                        const bool ok = ((snib >> i) & 1u) && tx > -1.0f && tx < Wf && ty > -1.0f && ty < Hf;
                        const unsigned bx = __float_as_uint(__fadd_rz(fmaxf(tx, 0.0f), 8388608.0f));
                        const unsigned by = __float_as_uint(__fadd_rz(fmaxf(ty, 0.0f), 8388608.0f));
                        if (ok) dst_t[by * uW + bx - sc_bias] = sc_val;
                    }
                    const float d = comp(Dc, i);
                    const float xh = fmaf((float)i, inv_fx, xh0);
                    const float ia = rcp_approx(d);
                    float l1[5], l2[5];
                    l1[0] = ia; l1[1] = -xh * ia; l1[2] = -xh * yh; l1[3] = fmaf(xh, xh, 1.0f); l1[4] = -yh;
                    l2[0] = ia; l2[1] = -yh * ia; l2[2] = -fmaf(yh, yh, 1.0f); l2[3] = xh * yh; l2[4] = xh;
                    bool valid;
                    float nr;
                    if (reuse) {
                        // pass A already applied the gates and computed the innovation norm of this pixel
                        nr = comp(Nc, i);
                        valid = cand && nr >= 0.f;
                    } else {
                        // hpp:252 gates
                        valid = cand && fabsf(dx) < 1e9f && fabsf(dy) < 1e9f && d > 0.f && d < max_d;
                        // predicted flow of both rows at once (packed FP32x2): p = sum_k (l1[k], l2[k]) * (xa[k], xb[k])
                        float2 pp = __fmul2_rn(make_float2(l1[0], l2[0]), make_float2(x[0], x[1]));
                        pp = __ffma2_rn(make_float2(l1[1], l2[1]), make_float2(x[2], x[2]), pp);
                        pp = __ffma2_rn(make_float2(l1[2], l2[2]), make_float2(x[3], x[3]), pp);
                        pp = __ffma2_rn(make_float2(l1[3], l2[3]), make_float2(x[4], x[4]), pp);
                        pp = __ffma2_rn(make_float2(l1[4], l2[4]), make_float2(x[5], x[5]), pp);
                        const float2 nn = __ffma2_rn(make_float2(-c1, -c2), pp, make_float2(dx, dy));
                        nr = sqrt_approx(fmaf(nn.x, nn.x, nn.y * nn.y));
                    }
                    if (PASS == 0) {
                        if (g.stride > 1) {
                            if (cand) norms_t[nslot++] = valid ? nr : -1.0f;  // compact: slot = rank / stride
                        } else if (valid) {
                            if (i == 0) nv.x = nr; else if (i == 1) nv.y = nr; else if (i == 2) nv.z = nr; else nv.w = nr;
                        }
                    } else if (kPacked) {
                        // branch-free: an invalid pixel contributes with weight 0 (its inputs are sanitised first)
                        float l = 1.0f;
                        if (wp.use) l = fmaxf(wp.coef * __expf(-fabsf(nr - wp.m) * wp.inv_b), 1e-6f) * wp.inv_lmax;
                        l = valid ? l : 0.f;
                        const float ias = valid ? ia : 0.f;
                        dx = valid ? dx : 0.f;
                        dy = valid ? dy : 0.f;
                        {
                            const float2 e[5] = {make_float2(ias, ias), make_float2(-xh * ias, -yh * ias), make_float2(l1[2], l2[2]),
                                                 make_float2(l1[3], l2[3]), make_float2(l1[4], l2[4])};
                            const float2 ll = make_float2(l, l);
                            float2 w[5];
#pragma unroll
                            for (int k = 0; k < 5; ++k) w[k] = __fmul2_rn(ll, e[k]);
                            int o = 0;
#pragma unroll
                            for (int r = 0; r < 5; ++r)
#pragma unroll
                                for (int q = r; q < 5; ++q) {
                                    acc2[o] = __ffma2_rn(w[r], e[q], acc2[o]);
                                    ++o;
                                }
                            const float2 zz = make_float2(dx, dy);
#pragma unroll
                            for (int k = 0; k < 5; ++k) acc2[15 + k] = __ffma2_rn(w[k], zz, acc2[15 + k]);
                            cnt_acc += valid ? 1.0f : 0.f;
                        }
                    } else if (valid) {
                        float l = 1.0f;
                        if (wp.use) l = fmaxf(wp.coef * __expf(-fabsf(nr - wp.m) * wp.inv_b), 1e-6f) * wp.inv_lmax;
                        AT e1[5], e2[5];
                        if (sizeof(AT) == 8) {
                            // FP64 per-pixel terms: 1/d from the FP32 approximation (rel. error < 2^-22) by two Newton steps
                            // (error^2 per step: 6e-14, then below 1 ulp)
                            const double xhd = fma((double)i, a.inv_fxd, xh0d);
                            const double dd = (double)d;
                            double r = (double)ia;
                            r = fma(r, fma(-dd, r, 1.0), r);
                            r = fma(r, fma(-dd, r, 1.0), r);
                            e1[0] = (AT)r; e1[1] = (AT)(-xhd * r); e1[2] = (AT)(-xhd * yhd); e1[3] = (AT)(1.0 + xhd * xhd); e1[4] = (AT)(-yhd);
                            e2[0] = (AT)r; e2[1] = (AT)(-yhd * r); e2[2] = (AT)(-(1.0 + yhd * yhd)); e2[3] = (AT)(xhd * yhd); e2[4] = (AT)xhd;
                        } else {
#pragma unroll
                            for (int k = 0; k < 5; ++k) {
                                e1[k] = (AT)l1[k];
                                e2[k] = (AT)l2[k];
                            }
                        }
                        AT w1[5], w2[5];
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            w1[k] = (AT)l * e1[k];
                            w2[k] = (AT)l * e2[k];
                        }
                        int o = 0;
#pragma unroll
                        for (int r = 0; r < 5; ++r)
#pragma unroll
                            for (int s = r; s < 5; ++s) {
                                acc[o] = fma(w1[r], e1[s], acc[o]);
                                acc[15 + o] = fma(w2[r], e2[s], acc[15 + o]);
                                ++o;
                            }
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            acc[30 + k] = fma(w1[k], (AT)dx, acc[30 + k]);
                            acc[35 + k] = fma(w2[k], (AT)dy, acc[35 + k]);
                        }
                        acc[40] += (AT)1;
                    }
                }
            }
            // pass A, stride 1: position-addressed norm slots, one coalesced 128-bit store per lane
            if (PASS == 0 && g.stride == 1) reinterpret_cast<float4*>(norms_t)[(g0 + j) * 32 + lane] = nv;
            return true;
        };
#pragma unroll 1
        for (int j = 0; j < 4; j += 2) {
            if (!unit_step(j, DA, F0A, F1A, NA, DB, F0B, F1B, NB)) break;
            if (!unit_step(j + 1, DB, F0B, F1B, NB, DA, F0A, F1A, NA)) break;
        }
        my_unit = nx_unit;
        nx_unit = nx2_unit;
        nx2_unit = -1;
    }

    if (kPacked) {
#pragma unroll
        for (int o = 0; o < 15; ++o) {
            acc[o] = (AT)acc2[o].x;
            acc[15 + o] = (AT)acc2[o].y;
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            acc[30 + k] = (AT)acc2[15 + k].x;
            acc[35 + k] = (AT)acc2[15 + k].y;
        }
        acc[40] = (AT)cnt_acc;
    }
    if (PASS == 1) {
        // one partial per WARP (no block barrier: warps finish at different times)
        double* out = a.partials + ((long long)t * a.max_blocks + blockIdx.x * (kThreads / 32) + warp) * kNAcc;
#pragma unroll
        for (int i = 0; i < kNAcc; ++i) {
            const AT s = warp_sum(acc[i]);
            if (lane == 0) out[i] = (double)s;
        }
    }
}

// ---- the streaming pass, TMA-staged ------------------------------------------------------------
// Same arithmetic as k_flow_pass<.., FAST = true, float, ..> at stride 1 with Laplacian weighting, but the operands
// reach the SM through the bulk-copy engine instead of through registers: every warp owns a ring of kRingStages
// shared-memory stages, one 128-px unit each (depth 512 B | flow 1 KiB | mask 128 B in pass A, pass-A norms 512 B in
// pass B).  Three lanes issue one cp.async.bulk each per unit (completion on the stage's mbarrier), kRingStages units
// ahead of the one being processed, so each warp keeps ~6-8 KB in flight without holding a single register for it -
// the register-prefetch version above has 1.5-2 KB per warp in flight and was latency-bound (DRAM 45 % / 31 %, issue
// 65 % / 54 %).  A warp takes a CONTIGUOUS chunk of its track's unit list; norm slots stay addressed by list position.
constexpr int kRingStages = 4;
template <int PASS>
struct RingLayout {
    static constexpr int kD = 0, kF = 512, kX = 1536;             // X: mask (pass A) / norms (pass B)
    static constexpr int kStage = kX + (PASS == 0 ? 128 : 512);  // bytes per stage (multiple of 16)
    static constexpr int kSmem = (kThreads / 32) * kRingStages * (kStage + 8);
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

template <int PASS, bool SCATTER>
__global__ void __launch_bounds__(kThreads, PASS == 0 ? 3 : 2) k_flow_pass_ring(PassArgs a) {
    extern __shared__ __align__(128) unsigned char ring_smem[];
    using L = RingLayout<PASS>;
    const int t = blockIdx.y;
    const VelCtl c = a.ctl[t];
    bool do_sc = false;
    uint8_t sc_val = 0;
    if (SCATTER) {
        const WarpPlan& p = a.plan[t];
        do_sc = p.fused != 0;
        sc_val = (uint8_t)p.uniform_val;
    }
    if (!c.enable && !do_sc) return;
    if (PASS == 1) {
        if (!c.enable) return;
        if (a.auto_threshold >= 0) {
            const bool small = a.wt_n[gridDim.y + t] < a.auto_threshold;
            if (small != (a.auto_take_small != 0)) return;  // the FP64 variant handles this track
        }
    }
    const Geom& g = a.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const char* mask_t = reinterpret_cast<const char*>(a.seg + (long long)t * a.seg_stride);
    const char* depth_t = reinterpret_cast<const char*>(a.ft.depth[c.prev_slot] + (long long)t * a.ft.depth_stride);
    const char* flow_t = reinterpret_cast<const char*>(a.ft.flow[c.cur_slot]) + (long long)t * a.ft.flow_stride * 4;
    float* norms_t = a.norms + (long long)t * a.norm_stride;
    const int HW = g.HW, W = g.W;
    const unsigned uW = (unsigned)g.W;
    uint8_t* dst_t = SCATTER ? a.state_dst + (long long)t * g.HW : nullptr;
    const float inv_fx = g.inv_fx, max_d = g.max_depth_f;
    const float inv_w = 1.0f / (float)g.W, Wf = (float)g.W, Hf = (float)g.H;
    const bool small_hw = g.HW < (1 << 24);
    const unsigned sc_bias = 0x4b000000u * (uW + 1u);
    const uint32_t thr4 = (uint32_t)a.thr * 0x01010101u;

    float x[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = (float)a.x_pred[(long long)t * a.x_stride + i];
    const float c1 = (float)(a.fx * c.dt), c2 = (float)(a.fy * c.dt);
    WeightParams wp;
    wp.use = 0;
    if (PASS == 1) wp = a.wp[t];
    const float w_k2 = -wp.inv_b * 1.4426950408889634f, w_cl = wp.coef * wp.inv_lmax, w_floor = 1e-6f * wp.inv_lmax;
    float2 acc2[PASS == 1 ? 20 : 1];
    float cnt_acc = 0.f;
#pragma unroll
    for (int i = 0; i < (PASS == 1 ? 20 : 1); ++i) acc2[i] = make_float2(0.f, 0.f);

    // this warp's ring and barriers
    unsigned char* my_ring = ring_smem + warp * kRingStages * L::kStage;
    const uint32_t ring_s = smem_u32(my_ring);
    const uint32_t bar_s = smem_u32(ring_smem + (kThreads / 32) * kRingStages * L::kStage) + warp * kRingStages * 8;
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < kRingStages; ++st) mbar_init(bar_s + 8 * st, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    // contiguous chunk of the track's list of non-empty units
    const int n_list = a.wt_n[t];
    const int n_warps = gridDim.x * (kThreads / 32);
    const int chunk = (n_list + n_warps - 1) / n_warps;
    const int k0 = (blockIdx.x * (kThreads / 32) + warp) * chunk;
    const int cnt = max(0, min(n_list, k0 + chunk) - k0);
    const int32_t* list = a.wt_list + (long long)t * a.n_units + k0;
    // unit ids: two windows of 32 list entries, one per lane each, refilled a window ahead
    int idsA = lane < cnt ? list[lane] : -1;
    int idsB = 32 + lane < cnt ? list[32 + lane] : -1;
    auto unit_of = [&](int k) { return __shfl_sync(0xffffffffu, (k & 32) ? idsB : idsA, k & 31); };
    // copy roles: lane 0 depth, lane 1 flow, lane 2 mask (pass A) / pass-A norms (pass B, addressed by list position)
    const char* cp_base = lane == 0 ? depth_t : lane == 1 ? flow_t
                          : PASS == 0 ? mask_t : reinterpret_cast<const char*>(norms_t + (long long)k0 * kUnitPx);
    const uint32_t cp_bpp = lane == 0 ? 4u : lane == 1 ? 8u : PASS == 0 ? 1u : 4u;  // bytes per pixel
    const uint32_t cp_off = lane == 0 ? L::kD : lane == 1 ? L::kF : L::kX;
    const bool cp_by_pos = PASS == 1 && lane == 2;
    constexpr uint32_t kTxPerPx = PASS == 0 ? 13u : 16u;
    auto issue = [&](int kp) {
        const int unit = unit_of(kp);
        const int st = kp % kRingStages;
        const uint32_t bar = bar_s + 8 * st;
        const uint32_t npx = (uint32_t)min(kUnitPx, HW - unit * kUnitPx);  // multiple of 16 (checked by the launcher)
        if (lane == 0) mbar_expect_tx(bar, npx * kTxPerPx);
        __syncwarp();
        if (lane < 3)
            bulk_g2s(ring_s + st * L::kStage + cp_off, cp_base + (long long)(cp_by_pos ? kp : unit) * (long long)(kUnitPx * cp_bpp),
                     npx * cp_bpp, bar);
    };
#pragma unroll 1
    for (int kp = 0; kp < min(kRingStages, cnt); ++kp) issue(kp);

#pragma unroll 1
    for (int k = 0; k < cnt; ++k) {
        if ((k & 31) == 0 && k > 0) {  // the window that just ran out gets the entries two windows ahead
            const int v = k + 32 + lane < cnt ? list[k + 32 + lane] : -1;
            if (k & 32) idsA = v; else idsB = v;
        }
        const int st = k % kRingStages;
        mbar_wait(bar_s + 8 * st, (uint32_t)(k / kRingStages) & 1u);
        const unsigned char* sp = my_ring + st * L::kStage;
        const float4 Dc = *reinterpret_cast<const float4*>(sp + L::kD + lane * 16);
        const float4 F0c = *reinterpret_cast<const float4*>(sp + L::kF + lane * 32);
        const float4 F1c = *reinterpret_cast<const float4*>(sp + L::kF + lane * 32 + 16);
        const int unit = unit_of(k);
        const int q = unit * 32 + lane;
        const bool in_plane = (q << 2) < HW;
        uint32_t nib = 0, snib = 0;
        float4 Nc = make_float4(-1.f, -1.f, -1.f, -1.f);
        if (PASS == 0) {
            const uint32_t m = in_plane ? *reinterpret_cast<const uint32_t*>(sp + L::kX + lane * 4) : 0u;
            nib = c.enable ? nibble_of(__vcmpgtu4(m, thr4)) : 0u;
            if (SCATTER && do_sc) {
                snib = nibble_of(__vcmpne4(m, 0u));
                if (q == 0) snib &= ~1u;  // mask_(0,0) = 0 (hpp:224)
            }
        } else {
            Nc = *reinterpret_cast<const float4*>(sp + L::kX + lane * 16);
            nib = in_plane ? ((Nc.x >= 0.f ? 1u : 0u) | (Nc.y >= 0.f ? 2u : 0u) | (Nc.z >= 0.f ? 4u : 0u) | (Nc.w >= 0.f ? 8u : 0u)) : 0u;
        }
        float4 nv = make_float4(-1.f, -1.f, -1.f, -1.f);
        if ((nib | snib) != 0u) {
            const int px = q << 2;
            int v, u0;
            if (small_hw) {
                v = (int)((float)px * inv_w);
                u0 = px - v * W;
                if (u0 < 0) { u0 += W; --v; }
                if (u0 >= W) { u0 -= W; ++v; }
            } else {
                v = px / W;
                u0 = px - v * W;
            }
            const float vf = (float)v, u0f = (float)u0;
            float yh, xh0;
            if (PASS == 1) {
                yh = (float)(((double)v - a.cyd) * a.inv_fyd);
                xh0 = (float)(((double)u0 - a.cxd) * a.inv_fxd);
            } else {
                yh = (vf - g.cy) * g.inv_fy;
                xh0 = (u0f - g.cx) * inv_fx;
            }
            // pass A: predicted flow H x^- as a polynomial in x^ with per-quad coefficients (y^ is constant over the quad):
            //   p1 = a (x0 - x2 x^) + (x4 - y^ x5) + x^ (-y^ x3 + x^ x4)      p2 = a (x1 - y^ x2) - (1 + y^2) x3 + x^ (y^ x4 + x5)
            const float pk0 = fmaf(-yh, x[5], x[4]), pk1 = -yh * x[3];
            const float pq0 = fmaf(-yh, x[2], x[1]), pq1 = -fmaf(yh, yh, 1.0f) * x[3], pq2 = fmaf(yh, x[4], x[5]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool cand = (nib >> i) & 1u;
                const float4 f = i < 2 ? F0c : F1c;
                float dx = (i & 1) ? f.z : f.x;
                float dy = (i & 1) ? f.w : f.y;
                if (SCATTER) {
                    const float tx = __fadd_rn(u0f + (float)i, dx), ty = __fadd_rn(vf, dy);
                    const bool ok = ((snib >> i) & 1u) && tx > -1.0f && tx < Wf && ty > -1.0f && ty < Hf;
                    const unsigned bx = __float_as_uint(__fadd_rz(fmaxf(tx, 0.0f), 8388608.0f));
                    const unsigned by = __float_as_uint(__fadd_rz(fmaxf(ty, 0.0f), 8388608.0f));
                    if (ok) dst_t[by * uW + bx - sc_bias] = sc_val;
                }
                const float d = comp(Dc, i);
                const float xh = fmaf((float)i, inv_fx, xh0);
                const float ia = rcp_approx(d);
                if (PASS == 0) {
                    const bool valid = cand && fabsf(dx) < 1e9f && fabsf(dy) < 1e9f && d > 0.f && d < max_d;
                    const float p1 = fmaf(ia, fmaf(-x[2], xh, x[0]), fmaf(xh, fmaf(xh, x[4], pk1), pk0));
                    const float p2 = fmaf(ia, pq0, fmaf(xh, pq2, pq1));
                    const float n1 = fmaf(-c1, p1, dx), n2 = fmaf(-c2, p2, dy);
                    const float nr = sqrt_approx(fmaf(n1, n1, n2 * n2));
                    if (valid) {
                        if (i == 0) nv.x = nr; else if (i == 1) nv.y = nr; else if (i == 2) nv.z = nr; else nv.w = nr;
                    }
                } else {
                    float l1[5], l2[5];
                    l1[0] = ia; l1[1] = -xh * ia; l1[2] = -xh * yh; l1[3] = fmaf(xh, xh, 1.0f); l1[4] = -yh;
                    l2[0] = ia; l2[1] = -yh * ia; l2[2] = -fmaf(yh, yh, 1.0f); l2[3] = xh * yh; l2[4] = xh;
                    const float nr = comp(Nc, i);
                    const bool valid = cand;
                    // max(coef exp(-|n - m| / b), 1e-6) / lmax with the constants folded per track
                    float l = wp.use ? fmaxf(w_cl * ex2_approx(fabsf(nr - wp.m) * w_k2), w_floor) : 1.0f;
                    l = valid ? l : 0.f;
                    const float ias = valid ? ia : 0.f;
                    dx = valid ? dx : 0.f;
                    dy = valid ? dy : 0.f;
                    const float2 e[5] = {make_float2(ias, ias), make_float2(-xh * ias, -yh * ias), make_float2(l1[2], l2[2]),
                                         make_float2(l1[3], l2[3]), make_float2(l1[4], l2[4])};
                    const float2 ll = make_float2(l, l);
                    float2 w[5];
#pragma unroll
                    for (int kk = 0; kk < 5; ++kk) w[kk] = __fmul2_rn(ll, e[kk]);
                    int o = 0;
#pragma unroll
                    for (int r = 0; r < 5; ++r)
#pragma unroll
                        for (int qq = r; qq < 5; ++qq) {
                            acc2[o] = __ffma2_rn(w[r], e[qq], acc2[o]);
                            ++o;
                        }
                    const float2 zz = make_float2(dx, dy);
#pragma unroll
                    for (int kk = 0; kk < 5; ++kk) acc2[15 + kk] = __ffma2_rn(w[kk], zz, acc2[15 + kk]);
                }
            }
            if (PASS == 1) cnt_acc += (float)__popc(nib);
        }
        if (PASS == 0 && c.enable) reinterpret_cast<float4*>(norms_t)[(long long)(k0 + k) * 32 + lane] = nv;
        // every lane has consumed its part of the stage (the values above were used): hand it back to the copy engine
        __syncwarp();
        if (k + kRingStages < cnt) issue(k + kRingStages);
    }

    if (PASS == 1) {
        double* out = a.partials + ((long long)t * a.max_blocks + blockIdx.x * (kThreads / 32) + warp) * kNAcc;
#pragma unroll
        for (int o = 0; o < 15; ++o) {
            const float s1 = warp_sum(acc2[o].x), s2 = warp_sum(acc2[o].y);
            if (lane == 0) {
                out[o] = (double)s1;
                out[15 + o] = (double)s2;
            }
        }
#pragma unroll
        for (int kk = 0; kk < 5; ++kk) {
            const float s1 = warp_sum(acc2[15 + kk].x), s2 = warp_sum(acc2[15 + kk].y);
            if (lane == 0) {
                out[30 + kk] = (double)s1;
                out[35 + kk] = (double)s2;
            }
        }
        const float sc = warp_sum(cnt_acc);
        if (lane == 0) out[40] = (double)sc;
    }
}

// ---- exact radix select of the upper median + Laplacian parameters ---------------------------
// The select runs as three data passes over the norm slots; the per-track bookkeeping that used to be separate
// one-block kernels (init, scan of the histogram, final parameters) is done by the LAST block of each pass to finish
// for that track (ticket counter + __threadfence), which removes five dependent launches from the critical path.
struct SelArgs {
    const float* norms; long long norm_stride;
    SelState* sel; uint32_t* hist; uint32_t* ticket;
    const VelCtl* ctl; const int32_t* wt_n; int n_tracks; int stride;
    WeightParams* wp;
};

// predicated shared-memory increments (compare + predicated reduction, no branch around the atomic)
__device__ __forceinline__ void red_shared_inc_if_eq(uint32_t smem_addr, uint32_t x, uint32_t y) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "setp.eq.u32 P1, %1, %2;\n"
        "@P1 red.shared.add.u32 [%0], 1;\n"
        "}\n" ::"r"(smem_addr),
        "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void red_shared_inc_if_nonneg(uint32_t smem_addr, uint32_t x) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "setp.ge.s32 P1, %1, 0;\n"
        "@P1 red.shared.add.u32 [%0], 1;\n"
        "}\n" ::"r"(smem_addr),
        "r"(x)
        : "memory");
}

__device__ __forceinline__ bool last_block_of_track(uint32_t* ticket, int t) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t v = atomicAdd(ticket + t, 1u);
        is_last = (v == gridDim.x - 1);
        if (is_last) ticket[t] = 0;
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

// histogram bin scan by the 256 threads of one block: finds the bin holding rank k, returns (bin, rank inside bin,
// total count); the global histogram is read past L1 and zeroed for the next use
template <int NB>
__device__ __forceinline__ void scan_bins(uint32_t* gh, uint32_t k_or_half, bool k_is_half, uint32_t& bin_out, uint32_t& k_out,
                                          uint32_t& total_out) {
    constexpr int PER = NB / kThreads;
    __shared__ uint32_t sh[kThreads / 32];
    __shared__ uint32_t res[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t loc[PER];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        loc[i] = __ldcg(gh + threadIdx.x * PER + i);
        gh[threadIdx.x * PER + i] = 0;
        sum += loc[i];
    }
    const uint32_t incl = (uint32_t)warp_scan_incl((int)sum, lane);
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, total = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
        if (w < warp) woff += sh[w];
        total += sh[w];
    }
    uint32_t before = woff + incl - sum;
    const uint32_t k = k_is_half ? (total >> 1) : k_or_half;
    if (threadIdx.x == 0) { res[0] = 0; res[1] = 0; res[2] = total; }
    __syncthreads();
    if (k >= before && k < before + sum) {  // exactly one thread (if total > 0)
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            if (k < before + loc[i]) {
                res[0] = threadIdx.x * PER + i;
                res[1] = k - before;
                break;
            }
            before += loc[i];
        }
    }
    __syncthreads();
    bin_out = res[0];
    k_out = res[1];
    total_out = res[2];
}

template <int LEVEL>
__global__ void __launch_bounds__(kThreads) k_sel_hist(SelArgs a) {
    const int t = blockIdx.y;
    SelState& st = a.sel[t];
    uint32_t n, prefix = 0;
    if (LEVEL == 0) {
        // norm slots written by pass A: 128 per listed unit (stride 1, gated-out / non-candidate slots hold -1) or one
        // per selected candidate (stride > 1)
        n = !a.ctl[t].enable ? 0u : a.stride > 1 ? (uint32_t)((a.wt_n[a.n_tracks + t] + a.stride - 1) / a.stride)
                                                 : (uint32_t)a.wt_n[t] * 128u;
        if (n == 0) {
            if (blockIdx.x == 0 && threadIdx.x == 0) { st.n = 0; st.n_entries = 0; }
            return;
        }
    } else {
        if (st.n == 0) return;
        n = st.n_entries;
        prefix = st.prefix;
    }
    const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = min(n, blockIdx.x * per);
    const uint32_t hi = min(n, lo + per);
    constexpr int NB = kSelBins;
    __shared__ uint32_t h[NB];
    for (int i = threadIdx.x; i < NB; i += kThreads) h[i] = 0;
    __syncthreads();
    const uint32_t* keys = reinterpret_cast<const uint32_t*>(a.norms + (long long)t * a.norm_stride);
    // 128-bit loads, two in flight per thread: a scalar loop keeps one 4-byte load in flight per thread, which is
    // ~10 % of the bytes in flight HBM3e needs (measured: 0.3 ms per pass instead of 0.06)
    // branch-free per key: one compare and a predicated shared-memory reduction (a gated-out slot, -1.0f, has the
    // sign bit set: its 12-bit bin is >= 2048 and never equals a prefix)
    const uint32_t h_s = (uint32_t)__cvta_generic_to_shared(h);
    const uint32_t pfx = prefix >> 20;
    auto count = [&](uint32_t key) {
        if (LEVEL == 0)
            red_shared_inc_if_nonneg(h_s + ((key >> 20) << 2), key);
        else
            red_shared_inc_if_eq(h_s + (((key >> 8) & 0xfffu) << 2), key >> 20, pfx);
    };
    const uint4* keys4 = reinterpret_cast<const uint4*>(keys);
    const uint32_t lo4 = (lo + 3) >> 2, hi4 = hi >> 2;  // whole uint4s inside [lo, hi)
    for (uint32_t i = lo + threadIdx.x; i < min(hi, lo4 << 2); i += kThreads) count(keys[i]);  // head
    uint32_t i4 = lo4 + threadIdx.x;
    for (; i4 + kThreads < hi4; i4 += 2 * kThreads) {
        const uint4 k0 = keys4[i4], k1 = keys4[i4 + kThreads];
        count(k0.x); count(k0.y); count(k0.z); count(k0.w);
        count(k1.x); count(k1.y); count(k1.z); count(k1.w);
    }
    if (i4 < hi4) {
        const uint4 k0 = keys4[i4];
        count(k0.x); count(k0.y); count(k0.z); count(k0.w);
    }
    if (hi4 >= lo4)
        for (uint32_t i = (hi4 << 2) + threadIdx.x; i < hi; i += kThreads) count(keys[i]);  // tail
    __syncthreads();
    uint32_t* gh = a.hist + (long long)t * kSelBins;
    for (int i = threadIdx.x; i < NB; i += kThreads)
        if (h[i]) atomicAdd(gh + i, h[i]);
    if (!last_block_of_track(a.ticket, t)) return;
    // ---- last block of this track: locate the bin of the upper median s[n/2] ----
    uint32_t bin, krem, total;
    scan_bins<NB>(gh, LEVEL == 0 ? 0u : st.k, LEVEL == 0, bin, krem, total);
    if (threadIdx.x == 0) {
        if (LEVEL == 0) {
            st.n = total;  // valid measurements
            st.n_entries = n;
            st.prefix = bin << 20;
            st.k = krem;
            st.pad2 = 0;
            st.less_cnt = 0;
            st.less_sum = 0.0;
            st.total_sum = 0.0;
            st.less_max_bits = 0;
        } else {
            st.prefix |= bin << 8;
            st.k = krem;
        }
    }
}

// Last pass over the norms, after two radix levels fixed the top 24 key bits of the upper median s[n/2]: histogram of
// the low 8 bits of the keys inside that 24-bit bin (all keys with the same low bits are the SAME float, so counts are
// enough to reconstruct sums there), and count / sum / max of everything below the bin, plus the grand total.
__device__ void sel_finish_warp(int t, int lane, SelState* __restrict__ sel, uint32_t* __restrict__ hist,
                                WeightParams* __restrict__ wp);

__global__ void __launch_bounds__(kThreads) k_sel_l2stats(SelArgs a) {
    const int t = blockIdx.y;
    SelState* __restrict__ sel = a.sel;
    uint32_t* __restrict__ hist = a.hist;
    if (sel[t].n == 0) {  // nothing to weight: publish neutral parameters
        if (blockIdx.x == 0 && threadIdx.x < 32) sel_finish_warp(t, threadIdx.x, sel, hist, a.wp);
        return;
    }
    const uint32_t n = sel[t].n_entries;
    const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = min(n, blockIdx.x * per);
    const uint32_t hi = min(n, lo + per);
    const uint32_t pbin = sel[t].prefix >> 8;
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t* keys = reinterpret_cast<const uint32_t*>(a.norms + (long long)t * a.norm_stride);
    float tot = 0.f, ls = 0.f, lm = 0.f;  // per-thread FP32 partials (<= a few hundred terms), FP64 across threads
    unsigned lc = 0;
    const uint32_t h_s = (uint32_t)__cvta_generic_to_shared(h);
    auto visit = [&](uint32_t key) {  // branch-free; gated-out slots (-1.0f) have the sign bit set: kb > pbin always
        const float v = __uint_as_float(key);
        const uint32_t kb = key >> 8;
        const bool below = kb < pbin;
        tot += (key >> 31) ? 0.f : v;
        ls += below ? v : 0.f;
        lc += below ? 1u : 0u;
        lm = fmaxf(lm, below ? v : 0.f);
        red_shared_inc_if_eq(h_s + ((key & 0xffu) << 2), kb, pbin);
    };
    const uint4* keys4 = reinterpret_cast<const uint4*>(keys);
    const uint32_t lo4 = (lo + 3) >> 2, hi4 = hi >> 2;
    for (uint32_t i = lo + threadIdx.x; i < min(hi, lo4 << 2); i += kThreads) visit(keys[i]);
    uint32_t i4 = lo4 + threadIdx.x;
    for (; i4 + kThreads < hi4; i4 += 2 * kThreads) {
        const uint4 k0 = keys4[i4], k1 = keys4[i4 + kThreads];
        visit(k0.x); visit(k0.y); visit(k0.z); visit(k0.w);
        visit(k1.x); visit(k1.y); visit(k1.z); visit(k1.w);
    }
    if (i4 < hi4) {
        const uint4 k0 = keys4[i4];
        visit(k0.x); visit(k0.y); visit(k0.z); visit(k0.w);
    }
    if (hi4 >= lo4)
        for (uint32_t i = (hi4 << 2) + threadIdx.x; i < hi; i += kThreads) visit(keys[i]);
    double dtot = warp_sum((double)tot), dls = warp_sum((double)ls);
    lc = (unsigned)warp_sum((int)lc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lm = fmaxf(lm, __shfl_xor_sync(0xffffffffu, lm, o));
    __shared__ double s_tot[kThreads / 32], s_ls[kThreads / 32];
    __shared__ unsigned s_lc[kThreads / 32];
    __shared__ float s_lm[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_tot[warp] = dtot;
        s_ls[warp] = dls;
        s_lc[warp] = lc;
        s_lm[warp] = lm;
    }
    __syncthreads();
    uint32_t* gh = hist + (long long)t * kSelBins;
    if (h[threadIdx.x]) atomicAdd(gh + threadIdx.x, h[threadIdx.x]);
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) {
            dtot += s_tot[w];
            dls += s_ls[w];
            lc += s_lc[w];
            lm = fmaxf(lm, s_lm[w]);
        }
        atomicAdd(&sel[t].total_sum, dtot);
        atomicAdd(&sel[t].less_sum, dls);
        atomicAdd(&sel[t].less_cnt, (unsigned long long)lc);
        atomicMax(&sel[t].less_max_bits, __float_as_uint(lm));
    }
    if (!last_block_of_track(a.ticket, t)) return;
    if (threadIdx.x < 32) sel_finish_warp(t, threadIdx.x, sel, hist, a.wp);
}

// one warp per track (the last block of k_sel_l2stats to finish): complete the select from the 256-bin histogram and
// emit the Laplacian parameters
__device__ void sel_finish_warp(int t, int lane, SelState* __restrict__ sel, uint32_t* __restrict__ hist,
                                WeightParams* __restrict__ wp) {
    SelState s;  // read past L1: the statistics were accumulated by other blocks' atomics
    {
        const uint4* p = reinterpret_cast<const uint4*>(sel + t);
        static_assert(sizeof(SelState) == 48, "SelState layout");
        const uint4 q0 = __ldcg(p), q1 = __ldcg(p + 1), q2 = __ldcg(p + 2);
        s.prefix = q0.x; s.k = q0.y; s.n = q0.z; s.n_entries = q0.w;
        s.less_cnt = (unsigned long long)q1.x | ((unsigned long long)q1.y << 32);
        s.less_sum = __hiloint2double((int)q1.w, (int)q1.z);
        s.total_sum = __hiloint2double((int)q2.y, (int)q2.x);
        s.less_max_bits = q2.z;
        s.pad2 = 0;
    }
    WeightParams w;
    w.m = 0.f;
    w.inv_b = 0.f;
    w.coef = 0.f;
    w.inv_lmax = 1.f;
    w.use = 0;
    w.n = (int32_t)s.n;
    w.pad[0] = w.pad[1] = 0;
    if (s.n > 0) {
        // bins of this lane: 8 consecutive low-byte values
        uint32_t* gh = hist + (long long)t * kSelBins;
        uint32_t loc[8];
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            loc[i] = __ldcg(gh + lane * 8 + i);
            gh[lane * 8 + i] = 0;
            sum += loc[i];
        }
        const uint32_t incl = (uint32_t)warp_scan_incl((int)sum, lane);
        uint32_t before = incl - sum;
        // the lane whose bins contain rank s.k decides the low byte of the upper median
        int bsel = -1;
        if (s.k >= before && s.k < before + sum) {
            uint32_t acc = before;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (bsel < 0 && s.k < acc + loc[i]) bsel = lane * 8 + i;
                acc += loc[i];
            }
        }
        const uint32_t vote = __ballot_sync(0xffffffffu, bsel >= 0);
        const int src = __ffs(vote) - 1;
        bsel = __shfl_sync(0xffffffffu, bsel, src < 0 ? 0 : src);
        // statistics of the in-bin keys below the selected one
        unsigned long long c_in = 0;
        double s_in = 0.0;
        float m_in = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int b = lane * 8 + i;
            if (b < bsel && loc[i]) {
                const float v = __uint_as_float((s.prefix & 0xffffff00u) | (uint32_t)b);
                c_in += loc[i];
                s_in += (double)loc[i] * (double)v;
                m_in = fmaxf(m_in, v);
            }
        }
        c_in = (unsigned long long)warp_sum((int)c_in);
        s_in = warp_sum(s_in);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m_in = fmaxf(m_in, __shfl_xor_sync(0xffffffffu, m_in, o));
        s.prefix = (s.prefix & 0xffffff00u) | (uint32_t)bsel;
        s.less_cnt += c_in;
        s.less_sum += s_in;
        const float less_max = fmaxf(__uint_as_float(s.less_max_bits), m_in);
        // SKFCorrection.cpp:95-102: median (even: mean of the two middle values), b = mean |n - m|
        const double n = (double)s.n;
        const double k1 = (double)(s.n >> 1);
        const double v1 = (double)__uint_as_float(s.prefix);
        const double lower = (s.less_cnt == (unsigned long long)(s.n >> 1)) ? (double)less_max : v1;
        const bool even = (s.n & 1u) == 0u;
        const double m = even ? 0.5 * (lower + v1) : v1;
        const double s_below = s.less_sum + (k1 - (double)s.less_cnt) * v1;  // sum of the k1 smallest
        const double s_above = s.total_sum - s_below;
        const double b = ((s_above - (n - k1) * m) + (k1 * m - s_below)) / n;
        if (b > 1e-4) {  // SKFCorrection.cpp:106
            const double dmin = even ? 0.5 * (v1 - lower) : 0.0;
            const double lmax = fmax(exp(-dmin / b) / (2.0 * b), 1e-6);
            w.m = (float)m;
            w.inv_b = (float)(1.0 / b);
            w.coef = (float)(1.0 / (2.0 * b));
            w.inv_lmax = (float)(1.0 / lmax);
            w.use = 1;
        }
    }
    if (lane == 0) wp[t] = w;
}

// ---- per-track epilogue: FP64 reduction of the block partials, 6x6 solve, gate, publish ------------
// Gauss-Jordan inverse of a symmetric positive definite 6x6 matrix (no pivoting needed for SPD input), by one warp:
// M = [A | I] in shared memory, on exit the right half holds A^-1.  Same elimination order as the scalar algorithm; the
// 72 entries of an elimination step are independent and spread over the lanes.
__device__ __noinline__ void spd6_inverse_warp(double (*M)[12], int lane) {
#pragma unroll 1
    for (int col = 0; col < 6; ++col) {
        const double piv = 1.0 / M[col][col];
        __syncwarp();
        if (lane < 12) M[col][lane] *= piv;
        __syncwarp();
        double f[3], m[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int e = lane + 32 * i, r = e / 12, c = e - r * 12;
            const bool on = e < 72 && r != col;
            f[i] = on ? M[r][col] : 0.0;
            m[i] = on ? M[col][c] : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int e = lane + 32 * i, r = e / 12, c = e - r * 12;
            if (e < 72 && r != col) M[r][c] -= f[i] * m[i];
        }
        __syncwarp();
    }
}

struct EpiArgs {
    int n_tracks;
    const VelCtl* ctl;
    const double* partials; int max_blocks; int n_blocks;
    double* v_mean; double* v_cov; const double* q_diag;
    double r0, r1, fx, fy;
    double* vel_hist; int hist_ring;
    int32_t* out_count; double* out_lambda; double* out_eta;
    int update_state;
};

__global__ void __launch_bounds__(32) k_vel_epilogue(EpiArgs a) {
    const int t = blockIdx.x;
    const int lane = threadIdx.x;
    const VelCtl c = a.ctl[t];
    __shared__ double sums[kNAcc];
    __shared__ double sLm[36], sEta[6], sRhs[6];
    __shared__ double M1[6][12], M2[6][12];
    __shared__ int s_count;
    if (c.enable) {
        // fixed summation order (deterministic): two interleaved chains per value keep the FP64 adds pipelined
        for (int i = lane; i < kNAcc; i += 32) {
            double s0 = 0.0, s1 = 0.0;
            const double* p = a.partials + (long long)t * a.max_blocks * kNAcc + i;
            int b = 0;
            for (; b + 1 < a.n_blocks; b += 2) {
                s0 += p[(long long)b * kNAcc];
                s1 += p[(long long)(b + 1) * kNAcc];
            }
            if (b < a.n_blocks) s0 += p[(long long)b * kNAcc];
            sums[i] = s0 + s1;
        }
    }
    __syncwarp();
    double* x = a.v_mean + (long long)t * 6;
    double* P = a.v_cov + (long long)t * 36;
    if (lane == 0) {
        int count = 0;
        for (int i = 0; i < 36; ++i) sLm[i] = 0.0;
        for (int i = 0; i < 6; ++i) sEta[i] = 0.0;
        if (c.enable) {
            const int i1[5] = {0, 2, 3, 4, 5}, i2[5] = {1, 2, 3, 4, 5};
            const double k1 = (a.fx * c.dt) * (a.fx * c.dt) / a.r0, k2 = (a.fy * c.dt) * (a.fy * c.dt) / a.r1;
            const double e1 = (a.fx * c.dt) / a.r0, e2 = (a.fy * c.dt) / a.r1;
            int o = 0;
            for (int r = 0; r < 5; ++r)
                for (int q = r; q < 5; ++q) {
                    const double v1 = k1 * sums[o], v2 = k2 * sums[15 + o];
                    sLm[i1[r] * 6 + i1[q]] += v1;
                    if (r != q) sLm[i1[q] * 6 + i1[r]] += v1;
                    sLm[i2[r] * 6 + i2[q]] += v2;
                    if (r != q) sLm[i2[q] * 6 + i2[r]] += v2;
                    ++o;
                }
            for (int k = 0; k < 5; ++k) {
                sEta[i1[k]] += e1 * sums[30 + k];
                sEta[i2[k]] += e2 * sums[35 + k];
            }
            count = (int)(sums[40] + 0.5);
        }
        s_count = count;
    }
    __syncwarp();
    const int count = s_count;
    if (a.out_lambda)
        for (int i = lane; i < 36; i += 32) a.out_lambda[(long long)t * 36 + i] = sLm[i];
    if (a.out_eta && lane < 6) a.out_eta[(long long)t * 6 + lane] = sEta[lane];
    // ROFTFilter.cpp:294-301: fewer than 3 valid pixels (or an empty measurement, SKFCorrection.cpp:60-68 keeps
    // the PREDICTED state, which the observability gate then reverts) -> the belief is left untouched.
    if (c.enable && a.update_state && count >= 3) {  // warp-uniform
        for (int e = lane; e < 72; e += 32) {
            const int r = e / 12, cc = e - r * 12;
            // KFPrediction: P + Q, F = I
            M1[r][cc] = cc < 6 ? P[r * 6 + cc] + (r == cc ? a.q_diag[r] : 0.0) : (cc - 6 == r ? 1.0 : 0.0);
        }
        __syncwarp();
        spd6_inverse_warp(M1, lane);  // right half: (P + Q)^-1
        for (int e = lane; e < 72; e += 32) {
            const int r = e / 12, cc = e - r * 12;
            M2[r][cc] = cc < 6 ? M1[r][6 + cc] + sLm[r * 6 + cc] : (cc - 6 == r ? 1.0 : 0.0);
        }
        __syncwarp();
        spd6_inverse_warp(M2, lane);  // right half: the corrected covariance
        if (lane < 6) {
            double v = sEta[lane];
            for (int j = 0; j < 6; ++j) v += M1[lane][6 + j] * x[j];
            sRhs[lane] = v;
        }
        __syncwarp();
        if (lane < 6) {
            double v = 0.0;
            for (int j = 0; j < 6; ++j) v += M2[lane][6 + j] * sRhs[j];
            x[lane] = v;
        }
        // symmetrise the information-form covariance (exactly symmetric in exact arithmetic)
        for (int e = lane; e < 36; e += 32) {
            const int i = e / 6, j = e - i * 6;
            P[e] = 0.5 * (M2[i][6 + j] + M2[j][6 + i]);
        }
    }
    __syncwarp();
    if (lane == 0 && a.out_count) a.out_count[t] = count;
    // velocity_->set_twist(v_corr_belief_.mean()) every frame (ROFTFilter.cpp:305)
    if (a.vel_hist && c.hist_slot >= 0 && lane < 6) {
        double* h = a.vel_hist + ((long long)t * a.hist_ring + c.hist_slot) * 6;
        h[lane] = x[lane];
    }
}

}  // namespace

int launch_mask_rank(const uint8_t* seg, long long seg_stride, int thr, int HW, int n_items, int32_t* wt_count, int32_t* total,
                     const VelCtl* ctl, cudaStream_t s) {
    const int n_warp_tiles = (HW + kWarpTilePx - 1) / kWarpTilePx;
    const int n_block_tiles = (HW + kBlockTilePx - 1) / kBlockTilePx;
    ROFTB_LAUNCH(k_mask_count, dim3(min(n_block_tiles, 64), n_items), kThreads, 0, s, seg, seg_stride, thr, HW, n_warp_tiles,
                 wt_count, ctl);
    ROFTB_LAUNCH(k_wt_scan, n_items, kThreads, 0, s, wt_count, n_warp_tiles, total, ctl);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_wt_scan(int32_t* wt_count, int n_warp_tiles, int n_items, int32_t* total, cudaStream_t s) {
    ROFTB_LAUNCH(k_wt_scan, n_items, kThreads, 0, s, wt_count, n_warp_tiles, total, (const VelCtl*)nullptr);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_tile_list(const uint8_t* plane, long long stride, int thr, int HW, int n_items, int32_t* wt_count, int32_t* wt_list,
                     int32_t* wt_n, const int32_t* active, int active_stride, cudaStream_t s) {
    const int n_units = (HW + kUnitPx - 1) / kUnitPx;
    const int n_block_tiles = (HW + kBlockTilePx - 1) / kBlockTilePx;
    const int bx = max(1, min(n_block_tiles, (148 * 8 + n_items - 1) / n_items));
    ROFTB_LAUNCH(k_tile_count, dim3(bx, n_items), kThreads, 0, s, plane, stride, thr, HW, n_units, wt_count, active,
                 active_stride);
    ROFTB_LAUNCH(k_tile_compact, n_items, kCompactThreads, 0, s, wt_count, wt_list, wt_n, n_units, active, active_stride);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_velocity(const VelocityArgs& a, cudaStream_t s) {
    const int T = a.n_tracks;
    const Geom& g = a.g;
    const int n_warp_tiles = (g.HW + kWarpTilePx - 1) / kWarpTilePx;
    const int n_block_tiles = (g.HW + kBlockTilePx - 1) / kBlockTilePx;
    // blocks per track: several waves over the machine in total; each warp walks the track's tile list with
    // a stride of (blocks x warps)
    int bpt = max(1, (148 * 16 + T - 1) / T);
    {
        static const int env_bpt = [] { const char* e = getenv("ROFTB_BPT"); return e ? atoi(e) : 0; }();
        if (env_bpt > 0) bpt = env_bpt;  // tuning hook
    }
    bpt = min(bpt, min(n_block_tiles, a.max_blocks / (kThreads / 32)));

    if (a.prof) cudaEventRecord(a.prof[1], s);
    PassArgs pa;
    pa.g = g;
    pa.ft = a.ft;
    pa.seg = a.seg;
    pa.seg_stride = a.seg_stride;
    pa.thr = a.thr;
    pa.ctl = a.ctl;
    pa.n_units = (g.HW + kUnitPx - 1) / kUnitPx;
    pa.norm_stride = (long long)pa.n_units * kUnitPx;
    pa.wt_prefix = a.wt_count;
    pa.wt_list = a.wt_list;
    pa.wt_n = a.wt_n;
    pa.norms = a.norms;
    pa.norm_count = a.norm_count;
    pa.wp = a.wp;
    pa.weight_flow = a.weight_flow;
    pa.x_pred = a.x_pred_override ? a.x_pred_override : a.v_mean;
    pa.x_stride = 6;
    pa.partials = a.partials;
    pa.max_blocks = a.max_blocks;
    pa.fx = a.fx;
    pa.fy = a.fy;
    pa.cxd = a.cx;
    pa.cyd = a.cy;
    pa.inv_fxd = 1.0 / a.fx;
    pa.inv_fyd = 1.0 / a.fy;
    pa.plan = a.plan;
    pa.state_dst = a.state_dst;
    pa.winner = a.winner;
    const bool fast = (!g.flow_s16 && g.grid == 1 && g.scale_mode == 0);
    const bool fuse = a.fuse_scatter != 0;  // the first streaming pass also propagates the mask
    cudaStream_t ps = s;                    // stream the ROFTB_PASS launches go to
#define ROFTB_PASS(PASS, AT, SC)                                                                          \
    do {                                                                                                  \
        if (fast)                                                                                         \
            ROFTB_LAUNCH((k_flow_pass<PASS, true, AT, SC>), dim3(bpt, T), kThreads, 0, ps, pa);           \
        else                                                                                              \
            ROFTB_LAUNCH((k_flow_pass<PASS, false, AT, SC>), dim3(bpt, T), kThreads, 0, ps, pa);          \
    } while (0)
    // TMA-staged variant of both passes for the common configuration (dense float2 flow, stride 1, weighting on)
    static const int env_ring = [] { const char* e = getenv("ROFTB_RING"); return e ? atoi(e) : 1; }();
    const bool ring = env_ring != 0 && fast && g.stride == 1 && a.weight_flow && (g.HW % 16) == 0;
    if (ring) {
        static bool attr_done = false;  // (per process; the attribute is per function)
        if (!attr_done) {
            cudaFuncSetAttribute(k_flow_pass_ring<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RingLayout<0>::kSmem);
            cudaFuncSetAttribute(k_flow_pass_ring<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RingLayout<0>::kSmem);
            cudaFuncSetAttribute(k_flow_pass_ring<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RingLayout<1>::kSmem);
            attr_done = true;
        }
    }
    if (a.weight_flow) {
        if (ring) {
            if (fuse)
                ROFTB_LAUNCH((k_flow_pass_ring<0, true>), dim3(bpt, T), kThreads, RingLayout<0>::kSmem, s, pa);
            else
                ROFTB_LAUNCH((k_flow_pass_ring<0, false>), dim3(bpt, T), kThreads, RingLayout<0>::kSmem, s, pa);
        } else if (fuse)
            ROFTB_PASS(0, float, true);
        else
            ROFTB_PASS(0, float, false);
        // the next step's preparation (prep stream) may start once the mask propagation of this pass is done; where it
        // is released decides which kernels of this step it shares the SMs with (ROFTB_PREP_AFTER: 0 = pass A,
        // 1 = first select pass, 2 = whole select - the default: the worklist kernels then overlap the issue-bound pass B
        // instead of the DRAM-bound select passes, measured 3-4 % faster per step)
        static const int prep_after = [] { const char* e = getenv("ROFTB_PREP_AFTER"); return e ? atoi(e) : 2; }();
        if (a.ev_first_pass && prep_after == 0) cudaEventRecord(a.ev_first_pass, s);
        if (a.prof) cudaEventRecord(a.prof[2], s);
        // enough blocks per track to spread the list, few enough that the per-block histogram flush stays cheap
        int sb = max(1, min(32, (148 * 24 + T - 1) / T));
        {
            static const int env_sb = [] { const char* e = getenv("ROFTB_SB"); return e ? atoi(e) : 0; }();
            if (env_sb > 0) sb = env_sb;  // tuning hook
        }
        SelArgs sa;
        sa.norms = a.norms; sa.norm_stride = pa.norm_stride; sa.sel = a.sel; sa.hist = a.hist; sa.ticket = a.norm_count;
        sa.ctl = a.ctl; sa.wt_n = a.wt_n; sa.n_tracks = T; sa.stride = g.stride; sa.wp = a.wp;
        ROFTB_LAUNCH(k_sel_hist<0>, dim3(sb, T), kThreads, 0, s, sa);
        if (a.ev_first_pass && prep_after == 1) cudaEventRecord(a.ev_first_pass, s);
        ROFTB_LAUNCH(k_sel_hist<1>, dim3(sb, T), kThreads, 0, s, sa);
        ROFTB_LAUNCH(k_sel_l2stats, dim3(sb, T), kThreads, 0, s, sa);
        if (a.ev_first_pass && prep_after >= 2) cudaEventRecord(a.ev_first_pass, s);
    } else if (a.prof) {
        cudaEventRecord(a.prof[2], s);
    }
    if (a.prof) cudaEventRecord(a.prof[3], s);
    const bool fuse_b = fuse && !a.weight_flow;
    // accum_fp64: 0 = FP32 terms, 1 = FP64 terms, 2 = auto: FP64 for tracks with fewer than kAutoFp64Candidates
    // candidate pixels (small, typically ill-conditioned problems where FP32 rounding is amplified most and FP64 costs
    // least), FP32 terms with FP64 reduction above (rounding averages out as 1/sqrt(N)); see DESIGN.md 4.1
    pa.auto_threshold = -1;
    pa.auto_take_small = 0;
    pa.reuse_norms = (a.weight_flow && g.stride == 1) ? 1 : 0;
    if (a.accum_fp64 == 2) {
        pa.auto_threshold = kAutoFp64Candidates;
        pa.auto_take_small = 1;
    }
    // auto: the FP64 launch only has the few small tracks to do (tens of microseconds of a mostly idle GPU) - it runs
    // on the auxiliary stream beside the FP32 launch (disjoint tracks, disjoint partial slots)
    const bool side = a.accum_fp64 == 2 && a.aux_stream && !fuse_b;
    if (a.accum_fp64 >= 1) {
        if (side) {
            cudaEventRecord(a.aux_fork, s);
            cudaStreamWaitEvent(a.aux_stream, a.aux_fork, 0);
            ps = a.aux_stream;
        }
        if (fuse_b)
            ROFTB_PASS(1, double, true);
        else
            ROFTB_PASS(1, double, false);
        if (side) cudaEventRecord(a.aux_join, a.aux_stream);
        ps = s;
    }
    if (a.accum_fp64 == 2) pa.auto_take_small = 0;
    if (a.accum_fp64 != 1) {
        if (ring)
            ROFTB_LAUNCH((k_flow_pass_ring<1, false>), dim3(bpt, T), kThreads, RingLayout<1>::kSmem, s, pa);
        else if (fuse_b)
            ROFTB_PASS(1, float, true);
        else
            ROFTB_PASS(1, float, false);
    }
#undef ROFTB_PASS
    if (side) cudaStreamWaitEvent(s, a.aux_join, 0);
    if (a.ev_first_pass && !a.weight_flow) cudaEventRecord(a.ev_first_pass, s);
    if (a.prof) cudaEventRecord(a.prof[4], s);
    EpiArgs e;
    e.n_tracks = T;
    e.ctl = a.ctl;
    e.partials = a.partials;
    e.max_blocks = a.max_blocks;
    e.n_blocks = bpt * (kThreads / 32);  // one partial per warp
    e.v_mean = a.v_mean;
    e.v_cov = a.v_cov;
    e.q_diag = a.q_diag;
    e.r0 = a.r_flow[0];
    e.r1 = a.r_flow[1];
    e.fx = a.fx;
    e.fy = a.fy;
    e.vel_hist = a.vel_hist;
    e.hist_ring = a.hist_ring;
    e.out_count = a.out_count;
    e.out_lambda = a.out_lambda;
    e.out_eta = a.out_eta;
    e.update_state = a.update_state;
    ROFTB_LAUNCH(k_vel_epilogue, T, 32, 0, s, e);
    if (a.prof) cudaEventRecord(a.prof[5], s);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace roftb
