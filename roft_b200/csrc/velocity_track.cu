// Optical-flow-aided velocity measurement + linear Kalman correction, batched over tracks
// (north-star part 1) - ONE kernel launch per step, one thread-block CLUSTER per track.
//
// Replaces, for every track at once:
//   ImageOpticalFlowMeasurement<T>::freeze        src/roft-lib/include/ROFT/ImageOpticalFlowMeasurement.hpp:231-283
//   SKFCorrection::correctStep                    src/roft-lib/src/SKFCorrection.cpp:37-153
//   SpatialVelocityModel + bfl::KFPrediction      src/roft-lib/src/SpatialVelocityModel.cpp:15-27
//   observability gate                            src/roft-lib/src/ROFTFilter.cpp:294-301
//   "no new mask" propagation of the mask sync    .../ImageSegmentationOFAidedSource.hpp:221-226 (fused into pass A)
//
// The reference materialises z (2N) and H (2N x 6) and then runs a SEQUENTIAL 2-row Kalman update per pixel.  For
// per-pixel independent noise that is algebraically the information-form sum
//     Lambda = (P+Q)^-1 + sum_j l_j H_j^T R^-1 H_j ,  eta = (P+Q)^-1 x + sum_j l_j H_j^T R^-1 z_j
// (SURVEY.md F1).  With x^ = (u-cx)/fx, y^ = (v-cy)/fy, a = 1/d the two rows of H_j are dt*fx*L1 and dt*fy*L2 with
//     L1 = [a, 0, -x^a, -x^y^, 1+x^2, -y^]     L2 = [0, a, -y^a, -(1+y^2), x^y^, x^]
// so the sums S1 = sum l L1^T L1, S2 = sum l L2^T L2, g1 = sum l L1^T nu1, g2 = sum l L2^T nu2 (nu = z - H x^-, the
// innovation) are a streaming reduction and a per-track FP64 epilogue solves the 6x6 system.
//
// The Laplacian weights l_j (SKFCorrection.cpp:91-116) need the median m and the mean absolute deviation b of N
// "norms" that the reference takes from a COLUMN-major view of the interleaved innovation vector (quirk Q3,
// DESIGN.md): r_i = sqrt(nu[i]^2 + nu[N+i]^2), i = 0..N-1, where nu = [nu1_0, nu2_0, nu1_1, ...] over the valid
// pixels in selection (row-major) order.  So the innovations have to exist in COMPACT selection order before the
// weights can: the work of a track is a chain of data-dependent phases
//     A   stream the listed units (mask, depth, flow): gates, innovations -> compact records; mask propagation
//     P   pair nu[i] with nu[N+i] -> r_i ; first radix level of the exact median select
//     S1,S2  remaining radix levels; the last one also gathers what b = mean|r - m| needs
//     B   stream the compact records: weights, 41 sums ; E  6x6 solve, gate, publish
// One cluster of C CTAs owns a track for all of them, separated by cluster barriers.  Everything a phase hands to the
// next (16 B per valid pixel + 4 B per r_i) is written and re-read by the same cluster within tens of microseconds
// and lives in a small pool of scratch slots that later tracks overwrite in place, so it stays in the 126 MB L2
// and never costs HBM traffic; clusters of different tracks are in different phases at any time, which overlaps the
// HBM-bound phase A of one track with the L2- / issue-bound phases of others - with a single launch per step.
#include <cooperative_groups.h>

#include <type_traits>

#include "roftb_internal.cuh"

namespace roftb {
namespace {

// ---- small PTX helpers ---------------------------------------------------------------------------
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
// same, with an L2 eviction-priority hint (the frame planes are read once: evict_first keeps them from displacing the
// scratch records the later phases re-read)
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}
// 16-byte asynchronous global -> shared copy of one lane (L1 bypassed), optionally with an L2 eviction-priority hint
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint64_t pol, int hint) {
    if (hint)
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
    else
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// barrier over every thread of the cluster; release / acquire at cluster scope orders the global-memory hand-over
// between the phases (the acquire side invalidates L1)
__device__ __forceinline__ void cluster_barrier() {
    __syncthreads();  // (also reconverges every warp for the .aligned barrier)
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned long long global_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void red_shared_inc(uint32_t smem_addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(smem_addr) : "memory");
}
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
// high bit of every non-zero byte of x
__device__ __forceinline__ uint32_t nz4(uint32_t x) { return (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u; }
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

constexpr int kVtWarps = kVtThreads / 32;
constexpr int kRingStages = 4;
constexpr int kStageD = 0, kStageF = 512, kStageM = 1536, kStageBytes = 1664;  // depth | flow | mask of one unit
constexpr int kRingBytes = kVtWarps * kRingStages * kStageBytes;
constexpr int kRingRegion = kRingBytes + kVtWarps * kRingStages * 8;            // + one mbarrier per stage
static_assert(kRingRegion >= kSelBins * 4, "the select histogram aliases the ring");

struct SelParams {   // Laplacian parameters of one track (SKFCorrection.cpp:95-116), folded for the per-pixel body
    float m;         // median of the Q3 norms
    float k2;        // -log2(e) / b
    float floor_;    // 2e-6 * b : the 1e-6 likelihood floor in units of exp(.)
    int use;         // b > 1e-4
    double b;
};

// exclusive scan of one int per thread across the block; total returned to every thread
__device__ __forceinline__ int block_excl_scan(int v, int* s_w, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int incl = warp_scan_incl(v, lane);
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kVtWarps; ++w) {
        const int x = s_w[w];
        if (w < warp) woff += x;
        tot += x;
    }
    __syncthreads();
    total = tot;
    return woff + incl - v;
}

// Gauss-Jordan inverse of a symmetric positive definite 6x6 matrix (no pivoting needed for SPD input), by one warp:
// M = [A | I] in shared memory, on exit the right half holds A^-1.
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

struct VtArgs {
    Geom g;
    FrameTable ft;
    const uint8_t* seg; long long seg_stride; int thr;   // candidate iff byte > thr
    const uint8_t* occ_src;                               // [T][n_units] occupancy flags of seg
    const VelCtl* ctl;
    int n_units;
    int weight_flow, accum_fp64, update_state;
    const double* x_pred;                                 // [T][6] mean the innovations are taken against
    double fx, fy, cx, cy, inv_fx, inv_fy;
    // fused single-flow mask propagation
    const WarpPlan* plan; uint8_t* state_dst; uint8_t* occ_dst;
    // scratch pool
    VelScratch sc;
    int32_t* wl_units; int32_t* wl_pixels;
    // scheduling order: cluster y processes track order[y] (largest worklist first, from the previous step's sizes);
    // the last cluster to finish writes the order of the next step
    const int32_t* order; int32_t* order_next; uint32_t* done_ticket;
    int first;                                            // this launch covers order[first .. first + gridDim.y)
    int total_tracks;                                     // tracks of the step over all launches (ticket target)
    unsigned long long* phase_clock;                      // optional [T][8]
    unsigned long long* span_clock;                       // optional [2]
    int l2_hint;
    int stage_mode;                                       // pass A staging: 0 = bulk copies + mbarrier, 1 = per-lane cp.async groups
    int slice_cap;                                        // list entries a CTA may own
    int smem_tab, smem_list;                              // byte offsets into dynamic shared memory
};

// position of pixel rank p (selection order over the whole track) inside the chunk-compacted record arrays:
// chunk c holds ranks [base[c], base[c+1]) at entries c * chunk_stride + (p - base[c]).  `c` is a lower bound hint.
__device__ __forceinline__ long long rec_index(int p, int& c, const int* s_base, long long chunk_stride) {
    while (p >= s_base[c + 1]) ++c;
    return (long long)c * chunk_stride + (p - s_base[c]);
}

// -------------------------------------------------------------------------------------------------
// Longest-processing-time-first order for the next step (tail of the launch = one small track, not one big one): a
// counting sort of the tracks by listed units, descending, by one warp.
__device__ void write_next_order(const VtArgs& a, int n_tracks, int lane, int* s_bucket /*[129]*/) {
    constexpr int NB = 128;
    for (int i = lane; i <= NB; i += 32) s_bucket[i] = 0;
    __syncwarp();
    const int shift = 32 - __clz(max(a.n_units, 1) / NB + 1);  // bucket = units >> shift < NB
    for (int t = lane; t < n_tracks; t += 32) {
        const int b = NB - 1 - min(NB - 1, __ldcg(a.wl_units + t) >> shift);  // descending
        atomicAdd(&s_bucket[b + 1], 1);
    }
    __syncwarp();
    if (lane == 0)
        for (int i = 1; i <= NB; ++i) s_bucket[i] += s_bucket[i - 1];
    __syncwarp();
    for (int t = lane; t < n_tracks; t += 32) {
        const int b = NB - 1 - min(NB - 1, __ldcg(a.wl_units + t) >> shift);
        a.order_next[atomicAdd(&s_bucket[b], 1)] = t;
    }
}

// REGS: register cap per thread.  96 (default) leaves a quarter of the register file - and, with ~75 KB of shared memory
// per CTA, a third of the shared memory - of every SM to the latency-bound kernels of the other streams (pose UKF,
// new-mask scatter) so that they run BESIDE the two resident CTAs of this kernel instead of waiting for it to drain;
// 80 allows three CTAs per SM, 128 is the uncapped build.
template <bool FAST, int REGS>
__global__ void __maxnreg__(REGS) k_velocity_track(const __grid_constant__ VtArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_w[kVtWarps];
    __shared__ int s_cnt[kVtMaxChunks];
    __shared__ int s_base[kVtMaxChunks + 1];
    __shared__ double s_part[kVtWarps][kVtPartN];
    __shared__ int s_misc[8];
    __shared__ SelParams s_sp;

    __shared__ int s_bucket[129];
    const int t = a.order ? a.order[a.first + blockIdx.y] : a.first + (int)blockIdx.y;
    const VelCtl c = a.ctl[t];
    bool do_sc = false;
    uint8_t sc_val = 0;
    if (a.plan) {
        const WarpPlan& p = a.plan[t];
        do_sc = p.fused != 0;  // only single-valued masks are fused (WarpPlan::fused)
        sc_val = (uint8_t)p.uniform_val;
    }
    const Geom& g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned rank = cluster_ctarank(), C = cluster_nctarank();
    // every cluster checks out once (rank 0, warp 0, all lanes converged); the last one orders the next step
    auto check_out = [&]() {
        if (!a.done_ticket) return;
        unsigned last = 0;
        if (lane == 0) {
            __threadfence();
            last = atomicAdd(a.done_ticket, 1u) == (unsigned)a.total_tracks - 1u ? 1u : 0u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            write_next_order(a, a.total_tracks, lane, s_bucket);
            if (lane == 0) *a.done_ticket = 0;
        }
    };
    if (!c.enable && !do_sc) {  // uniform over the cluster
        if (rank == 0 && warp == 0) check_out();
        return;
    }
    const int n_chunks = (int)C * kVtWarps;
    const int gchunk = (int)rank * kVtWarps + warp;
    const int HW = g.HW, W = g.W, nq = HW >> 2;
    const unsigned uW = (unsigned)W;
    unsigned long long* clk = (a.phase_clock && rank == 0 && tid == 0) ? a.phase_clock + (long long)t * 8 : nullptr;
    if (clk) clk[0] = global_timer();
    if (rank == 0) span_stamp(a.span_clock, false);
    // scratch slot: rank 0 claims one bit of the pool bitmap right away - the atomic's round trip overlaps the table and
    // worklist set-up below (the slot is published to the other CTAs at cluster barrier #1)
    if (c.enable && rank == 0 && tid == 0) {
        int sl = -1;
        while (sl < 0) {
            for (int w = 0; w < a.sc.n_slot_words && sl < 0; ++w) {
                uint32_t cur = *reinterpret_cast<volatile uint32_t*>(a.sc.slot_bitmap + w);
                while (~cur) {
                    const int bit = __ffs(~cur) - 1;
                    const uint32_t old = atomicOr(a.sc.slot_bitmap + w, 1u << bit);
                    if (!(old & (1u << bit))) { sl = w * 32 + bit; break; }
                    cur = old | (1u << bit);
                }
            }
        }
        s_misc[0] = sl;
        a.sc.track_slot[t] = sl;
    }
    // flags of the destination plane's previous content (lazy clear below): in flight during the worklist set-up
    uint4 pre_of = make_uint4(0u, 0u, 0u, 0u);
    const int pre_gi = (int)rank * kVtThreads + tid;
    bool pre_ok = false;
    if (do_sc) {
        const uint8_t* of0 = a.occ_dst + (long long)t * a.n_units;
        pre_ok = ((reinterpret_cast<uintptr_t>(of0) & 15) == 0) && pre_gi * 16 + 16 <= a.n_units;
        if (pre_ok) pre_of = __ldcg(reinterpret_cast<const uint4*>(of0) + pre_gi);
    }

    float* s_xh = reinterpret_cast<float*>(smem + a.smem_tab);
    float* s_yh = s_xh + W;
    int32_t* s_list = reinterpret_cast<int32_t*>(smem + a.smem_list);
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem);  // aliases the ring (used after pass A)

    // ================================ prologue =====================================================
    // normalised coordinates, correctly rounded from FP64 (a rounded 1/fx would bias every pixel the same way)
    {
        for (int i = tid; i < W; i += kVtThreads) s_xh[i] = (float)(((double)i - a.cx) * a.inv_fx);
        for (int i = tid; i < g.H; i += kVtThreads) s_yh[i] = (float)(((double)i - a.cy) * a.inv_fy);
    }
    // worklist: compact the occupancy flags of the mask plane; this CTA keeps the slice its warps will walk
    int n_list, per_chunk, slice_lo;
    {
        // flags staged in shared memory (the ring area is free until pass A) with coalesced 128-bit loads: one L2 round
        // trip instead of a serial byte loop; every thread then owns `per_t` (a multiple of 16) consecutive units
        const uint8_t* fl = a.occ_src + (long long)t * a.n_units;
        const int nu16 = (a.n_units + 15) >> 4;
        const bool staged = nu16 * 16 <= kRingRegion && ((reinterpret_cast<uintptr_t>(fl) & 15) == 0);
        if (staged) {
            uint4* sf = reinterpret_cast<uint4*>(smem);
            const uint4* gf = reinterpret_cast<const uint4*>(fl);
            for (int i = tid; i < nu16; i += kVtThreads) {
                uint4 w = make_uint4(0u, 0u, 0u, 0u);
                if (i * 16 + 16 <= a.n_units) {
                    w = __ldcg(gf + i);
                } else {  // ragged tail: byte by byte
                    uint32_t ws[4] = {0u, 0u, 0u, 0u};
                    for (int b = 0; b < 16; ++b)
                        if (i * 16 + b < a.n_units && fl[i * 16 + b]) ws[b >> 2] |= 1u << (8 * (b & 3));
                    w = make_uint4(ws[0], ws[1], ws[2], ws[3]);
                }
                sf[i] = w;
            }
            __syncthreads();
            fl = smem;
        }
        const int per_t = (((a.n_units + kVtThreads - 1) / kVtThreads) + 15) & ~15;
        const int u0 = min(a.n_units, tid * per_t), u1 = min(a.n_units, u0 + per_t);
        int cnt = 0;
        if (staged) {
            for (int u = u0; u < u1; u += 16) {  // (the staged copy is zero-padded to whole 16-byte groups)
                const uint4 w = *reinterpret_cast<const uint4*>(smem + u);
                cnt += (__popc(__vcmpne4(w.x, 0u)) + __popc(__vcmpne4(w.y, 0u)) + __popc(__vcmpne4(w.z, 0u)) + __popc(__vcmpne4(w.w, 0u))) >> 3;
            }
        } else {
            for (int u = u0; u < u1; ++u) cnt += fl[u] ? 1 : 0;
        }
        int pos = block_excl_scan(cnt, s_w, n_list);
        per_chunk = (n_list + n_chunks - 1) / n_chunks;
        slice_lo = (int)rank * kVtWarps * per_chunk;
        const int slice_hi = min(n_list, slice_lo + kVtWarps * per_chunk);
        if (pos < slice_hi && pos + cnt > slice_lo)  // this thread's units intersect the slice
            for (int u = u0; u < u1; ++u)
                if (fl[u]) {
                    if (pos >= slice_lo && pos < slice_hi) s_list[pos - slice_lo] = u;
                    ++pos;
                }
        __syncthreads();  // the staged flags are dead: the ring may be initialised
    }
    // lazy clear of the destination plane of the fused propagation: only the units its previous content occupied
    if (do_sc) {
        uint8_t* of = a.occ_dst + (long long)t * a.n_units;
        uint8_t* dst_t = a.state_dst + (long long)t * HW;
        // 16 flags per 128-bit load (the flag rows are 16-byte aligned whenever n_units is a multiple of 16)
        const bool vec = ((reinterpret_cast<uintptr_t>(of) & 15) == 0);
        const int ng = (a.n_units + 15) >> 4;
        for (int gi = (int)rank * kVtThreads + tid; gi < ng; gi += (int)C * kVtThreads) {
            const int ub = gi * 16;
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            if (vec && ub + 16 <= a.n_units) {
                const uint4 v4 = (pre_ok && gi == pre_gi) ? pre_of : *reinterpret_cast<const uint4*>(of + ub);
                w[0] = v4.x; w[1] = v4.y; w[2] = v4.z; w[3] = v4.w;
                if (v4.x | v4.y | v4.z | v4.w) *reinterpret_cast<uint4*>(of + ub) = make_uint4(0u, 0u, 0u, 0u);
            } else {
                for (int b = 0; b < 16 && ub + b < a.n_units; ++b)
                    if (of[ub + b]) {
                        w[b >> 2] |= 1u << (8 * (b & 3));
                        of[ub + b] = 0;
                    }
            }
#pragma unroll
            for (int b = 0; b < 16; ++b)
                if ((w[b >> 2] >> (8 * (b & 3))) & 0xffu) {
                    const int u = ub + b;
                    const int n16 = min(kUnitPx, HW - u * kUnitPx) >> 4;
                    uint4* p = reinterpret_cast<uint4*>(dst_t + (long long)u * kUnitPx);
                    for (int i = 0; i < n16; ++i) p[i] = make_uint4(0u, 0u, 0u, 0u);
                }
        }
    }
    // scratch slot (rank 0 claims one bit of the pool bitmap) and its histograms
    int slot = -1;
    if (c.enable) {
        if (rank == 0) {
            if (tid == 0) {  // (the slot itself was claimed at the top of the kernel)
                if (a.wl_units) a.wl_units[t] = n_list;
                if (a.wl_pixels) a.wl_pixels[t] = 0;
            }
            __syncthreads();
            slot = s_misc[0];  // (its histograms were left zeroed by the previous user)
        }
    }
    // stride > 1: candidates in every chunk, for the row-major rank (counted before the gates, hpp:237)
    const uint32_t thr4 = (uint32_t)a.thr * 0x01010101u;
    const uint32_t* mq = reinterpret_cast<const uint32_t*>(a.seg + (long long)t * a.seg_stride);
    const int my_k0 = gchunk * per_chunk;                          // first list position of this warp's chunk
    const int my_cnt = max(0, min(n_list, my_k0 + per_chunk) - my_k0);
    const int32_t* my_list = s_list + warp * per_chunk;
    __syncthreads();  // s_list, tables
    if (!FAST && g.stride > 1 && c.enable) {
        int run = 0;
        for (int k = 0; k < my_cnt; ++k) {
            const int q = my_list[k] * 32 + lane;
            const uint32_t m = q < nq ? ld_nc_u32(mq + q) : 0u;
            const int cu = __reduce_add_sync(0xffffffffu, __popc(__vcmpgtu4(m, thr4)) >> 3);
            run += cu;
        }
        if (lane == 0) a.sc.chunk_aux[(long long)t * kVtMaxChunks + gchunk] = run;
    }
    cluster_barrier();  // ---- #1: destination plane cleared, slot published, histograms zeroed
    if (clk) clk[1] = global_timer();
    if (c.enable && rank != 0) slot = __ldcg(a.sc.track_slot + t);
    int rank_base = 0;
    if (!FAST && g.stride > 1 && c.enable) {
        for (int i = 0; i < gchunk; ++i) rank_base += __ldcg(a.sc.chunk_aux + (long long)t * kVtMaxChunks + i);
    }

    // ================================ pass A =======================================================
    // Each warp walks a CONTIGUOUS chunk of the track's unit list and appends the valid measurements, in row-major
    // order, to its own region of the record arrays (region = list position x 128 entries): 4 ballots per unit give
    // every lane its offset, no atomics.  Records: nu = (nu1, nu2) | dp = (depth bits, v << 16 | u).
    const long long chunk_stride = (long long)per_chunk * kUnitPx;
    float2* nu_t = nullptr;
    uint2* dp_t = nullptr;
    float* r_t = nullptr;
    if (c.enable) {
        nu_t = a.sc.nu + (long long)slot * a.sc.cap;
        dp_t = a.sc.dp + (long long)slot * a.sc.cap;
        r_t = a.sc.r + (long long)slot * a.sc.cap;
    }
    {
        const char* depth_t = reinterpret_cast<const char*>(a.ft.depth[c.prev_slot] + (long long)t * a.ft.depth_stride);
        const char* fbase = reinterpret_cast<const char*>(a.ft.flow[c.cur_slot]) +
                            (long long)t * a.ft.flow_stride * (g.flow_s16 ? 2 : 4);
        const char* mask_t = reinterpret_cast<const char*>(mq);
        uint8_t* dst_t = a.state_dst ? a.state_dst + (long long)t * HW : nullptr;
        uint8_t* odst_t = a.occ_dst ? a.occ_dst + (long long)t * a.n_units : nullptr;
        const float max_d = g.max_depth_f;
        const float inv_w = 1.0f / (float)W, Wf = (float)W, Hf = (float)g.H;
        const bool small_hw = HW < (1 << 24);
        const unsigned sc_bias = 0x4b000000u * (uW + 1u);
        float x[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = (float)a.x_pred[(long long)t * 6 + i];
        const float c1 = (float)(a.fx * c.dt), c2 = (float)(a.fy * c.dt);
        const unsigned ustride = (unsigned)g.stride;
        const unsigned lt_mask = (1u << lane) - 1u;
        int woff = 0;       // records appended by this warp so far (warp-uniform)
        int cand_cnt = 0;   // candidates seen by this lane (diagnostics)
        int last_flag = -1; // last destination unit this lane flagged
        unsigned urank = (unsigned)rank_base;
        const long long rec0 = (long long)gchunk * chunk_stride;
        float2* nu_c = nu_t ? nu_t + rec0 : nullptr;   // this warp's region of the record arrays (32-bit offsets from here)
        uint2* dp_c = dp_t ? dp_t + rec0 : nullptr;
        const bool thr_simple = a.thr <= 1;
        const uint32_t thr_keep = a.thr == 1 ? 0xfefefefeu : 0xffffffffu;

        // ---- ring (FAST): depth 512 B | flow 1 KiB | mask 128 B per stage, three bulk copies per unit
        unsigned char* my_ring = smem + warp * kRingStages * kStageBytes;
        const uint32_t ring_s = smem_u32(my_ring);
        const uint32_t bar_s = smem_u32(smem + kRingBytes) + warp * kRingStages * 8;
        uint64_t pol = 0;
        if (FAST) {
            if (a.l2_hint) pol = policy_evict_first();
            if (lane == 0) {
#pragma unroll
                for (int st = 0; st < kRingStages; ++st) mbar_init(bar_s + 8 * st, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncwarp();
        }
        // one lane arms the stage's barrier and issues the three copies (the operands of a bulk copy live in uniform
        // registers: per-lane operands would be serialised lane by lane)
        auto issue = [&](int kp) {
            const int unit = my_list[kp];
            const int st = kp % kRingStages;
            const uint32_t bar = bar_s + 8 * st;
            const uint32_t npx = (uint32_t)min(kUnitPx, HW - unit * kUnitPx);  // multiple of 16 (checked by the launcher)
            const uint32_t dsts = ring_s + st * kStageBytes;
            const long long uoff = (long long)unit * kUnitPx;
            if (lane == 0) {
                mbar_expect_tx(bar, npx * 13u);
                if (a.l2_hint) {
                    bulk_g2s_hint(dsts + kStageD, depth_t + uoff * 4, npx * 4u, bar, pol);
                    bulk_g2s_hint(dsts + kStageF, fbase + uoff * 8, npx * 8u, bar, pol);
                    bulk_g2s_hint(dsts + kStageM, mask_t + uoff, npx, bar, pol);
                } else {
                    bulk_g2s(dsts + kStageD, depth_t + uoff * 4, npx * 4u, bar);
                    bulk_g2s(dsts + kStageF, fbase + uoff * 8, npx * 8u, bar);
                    bulk_g2s(dsts + kStageM, mask_t + uoff, npx, bar);
                }
            }
        };
        // the same staging with per-lane 16-byte asynchronous copies (cp.async): every lane issues four copies with
        // its own addresses - ~10 issue slots per unit, where the three bulk copies (whose operands must be made
        // uniform one copy at a time) cost ~70; one commit group per loop trip, empty past the end of the chunk
        const uint32_t l16 = (uint32_t)lane * 16u;
        auto issue_cpa = [&](int kp) {
            if (kp < my_cnt) {
                const int unit = my_list[kp];
                const uint32_t npx = (uint32_t)min(kUnitPx, HW - unit * kUnitPx);
                const uint32_t dsts = ring_s + (kp % kRingStages) * kStageBytes + l16;
                const long long uoff = (long long)unit * kUnitPx;
                if (l16 < npx * 4u) cp_async16(dsts + kStageD, depth_t + uoff * 4 + l16, pol, a.l2_hint);
                if (l16 < npx * 8u) cp_async16(dsts + kStageF, fbase + uoff * 8 + l16, pol, a.l2_hint);
                if (l16 + 512u < npx * 8u) cp_async16(dsts + kStageF + 512, fbase + uoff * 8 + 512 + l16, pol, a.l2_hint);
                if (l16 < npx) cp_async16(dsts + kStageM, mask_t + uoff + l16, pol, a.l2_hint);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        const bool cpa = a.stage_mode == 1;
        if (FAST) {
#pragma unroll 1
            for (int kp = 0; kp < kRingStages; ++kp) {
                if (cpa) issue_cpa(kp);
                else if (kp < my_cnt) issue(kp);
            }
        }
        uint32_t m_next = 0;
        if (!FAST && my_cnt > 0) {
            const int q = my_list[0] * 32 + lane;
            m_next = q < nq ? ld_nc_u32(mq + q) : 0u;
        }

#pragma unroll 1
        for (int k = 0; k < my_cnt; ++k) {
            const int unit = my_list[k];
            const int q = unit * 32 + lane;
            const bool in_plane = q < nq;
            float4 Dc = make_float4(0.f, 0.f, 0.f, 0.f), F0c = Dc, F1c = Dc;
            uint32_t m;
            if (FAST) {
                const int st = k % kRingStages;
                if (cpa) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(kRingStages - 1) : "memory");
                    __syncwarp();  // the lanes read each other's copies
                } else {
                    mbar_wait(bar_s + 8 * st, (uint32_t)(k / kRingStages) & 1u);
                }
                const unsigned char* sp = my_ring + st * kStageBytes;
                Dc = *reinterpret_cast<const float4*>(sp + kStageD + lane * 16);
                F0c = *reinterpret_cast<const float4*>(sp + kStageF + lane * 32);
                F1c = *reinterpret_cast<const float4*>(sp + kStageF + lane * 32 + 16);
                m = in_plane ? *reinterpret_cast<const uint32_t*>(sp + kStageM + lane * 4) : 0u;
            } else {
                m = m_next;
                if (k + 1 < my_cnt) {  // next unit's mask word in flight while this one is processed
                    const int qn = my_list[k + 1] * 32 + lane;
                    m_next = qn < nq ? ld_nc_u32(mq + qn) : 0u;
                }
            }
            // candidate / non-zero pixels of the quad as the high bit of each byte (no nibble packing, no per-byte
            // compare: byte > thr for thr in {0, 1} is "byte & ~thr != 0"; other thresholds take the compare)
            uint32_t tc = 0, ts = 0;
            if (c.enable) tc = thr_simple ? nz4(m & thr_keep) : (__vcmpgtu4(m, thr4) & 0x80808080u);
            if (do_sc) {
                ts = nz4(m);
                if (q == 0) ts &= ~0x80u;  // mask_(0,0) = 0 (hpp:224)
            }
            cand_cnt += __popc(tc);
            if (!FAST && g.stride > 1) {
                // keep rank % stride == 0 over the row-major rank of the candidates (hpp:237), BEFORE the gates
                const int cnt = __popc(tc);
                const int incl = warp_scan_incl(cnt, lane);
                unsigned rk = urank + (unsigned)(incl - cnt);
                urank += (unsigned)__shfl_sync(0xffffffffu, incl, 31);
                unsigned rem = rk % ustride;
                uint32_t ns = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (tc & (0x80u << (8 * i))) {
                        if (rem == 0u) ns |= 0x80u << (8 * i);
                        rem = rem + 1u == ustride ? 0u : rem + 1u;
                    }
                tc = ns;
            }
            bool valid[4] = {false, false, false, false};  // valid measurements of this lane's quad
            float n1[4], n2[4], dd[4];
            uint32_t pk0 = 0;
            if ((tc | ts) != 0u) {
                const int px = q << 2;
                int v, u0;
                if (small_hw) {  // reciprocal multiply with a +-1 fix-up (exact for HW < 2^24)
                    v = (int)((float)px * inv_w);
                    u0 = px - v * W;
                    if (u0 < 0) { u0 += W; --v; }
                    if (u0 >= W) { u0 -= W; ++v; }
                } else {
                    v = px / W;
                    u0 = px - v * W;
                }
                pk0 = ((uint32_t)v << 16) | (uint32_t)u0;
                const float vf = (float)v, u0f = (float)u0;
                const float yh = s_yh[v];
                const float4 xh4 = *reinterpret_cast<const float4*>(s_xh + u0);
                if (!FAST && tc) Dc = ld_nc_f4(reinterpret_cast<const float4*>(depth_t) + q);
                // predicted flow H x^- as a polynomial in x^ with per-quad coefficients (y^ is constant over the quad):
                //   p1 = a (x0 - x2 x^) + (x4 - y^ x5) + x^ (-y^ x3 + x^ x4)      p2 = a (x1 - y^ x2) - (1 + y^2) x3 + x^ (y^ x4 + x5)
                const float pk0c = fmaf(-yh, x[5], x[4]), pk1c = -yh * x[3];
                const float pq0 = fmaf(-yh, x[2], x[1]), pq1 = -fmaf(yh, yh, 1.0f) * x[3], pq2 = fmaf(yh, x[4], x[5]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool cand = (tc & (0x80u << (8 * i))) != 0u;
                    const bool scp = (ts & (0x80u << (8 * i))) != 0u;
                    float dx, dy;
                    if (FAST) {
                        const float4 f = i < 2 ? F0c : F1c;
                        dx = (i & 1) ? f.z : f.x;
                        dy = (i & 1) ? f.w : f.y;
                    } else {
                        dx = 0.f;
                        dy = 0.f;
                        if (cand || scp) {
                            const float2 f = load_flow(fbase, (long long)(v / g.grid) * g.Wf + ((u0 + i) / g.grid), g);
                            dx = f.x;
                            dy = f.y;
                        }
                    }
                    if (do_sc) {
                        // hpp:249-278 for a single flow: the source pixel is inside the frame and its flow element is the
                        // one just fetched; IEEE adds.  In-frame test on the floats: (int)t in [0, n) <=> -1 < t < n
                        // (truncation; NaN fails both; +-inf / huge values fail like x86's INT_MIN).  trunc(max(t, 0))
                        // without the conversion pipe: t + 2^23 rounded toward zero keeps floor(t) in the mantissa, so the
                        // float bits are 0x4b000000 + floor(t); the bias of both terms is folded into one constant.
                        const float tx = __fadd_rn(u0f + (float)i, dx), ty = __fadd_rn(vf, dy);
                        const bool ok = scp && tx > -1.0f && tx < Wf && ty > -1.0f && ty < Hf;
                        const unsigned bx = __float_as_uint(__fadd_rz(fmaxf(tx, 0.0f), 8388608.0f));
                        const unsigned by = __float_as_uint(__fadd_rz(fmaxf(ty, 0.0f), 8388608.0f));
                        const unsigned di = by * uW + bx - sc_bias;
                        if (ok) dst_t[di] = sc_val;
                        // occupancy flag of the destination unit (idempotent store, skipped while the unit repeats)
                        const int du = (int)(di >> 7);
                        const bool nf = ok && du != last_flag;
                        if (nf) odst_t[du] = 1;
                        last_flag = nf ? du : last_flag;
                    }
                    const float d = comp(Dc, i);
                    const float xh = comp(xh4, i);
                    const float ia = rcp_approx(d);
                    // hpp:252 gates (a NaN flow fails the magnitude tests)
                    valid[i] = cand && fabsf(dx) < 1e9f && fabsf(dy) < 1e9f && d > 0.f && d < max_d;
                    const float p1 = fmaf(ia, fmaf(-x[2], xh, x[0]), fmaf(xh, fmaf(xh, x[4], pk1c), pk0c));
                    const float p2 = fmaf(ia, pq0, fmaf(xh, pq2, pq1));
                    n1[i] = fmaf(-c1, p1, dx);
                    n2[i] = fmaf(-c2, p2, dy);
                    dd[i] = d;
                }
            }
            // in-unit compaction: pixel order = lane-major, then the four pixels of the quad
            const unsigned b0 = __ballot_sync(0xffffffffu, valid[0]), b1 = __ballot_sync(0xffffffffu, valid[1]);
            const unsigned b2 = __ballot_sync(0xffffffffu, valid[2]), b3 = __ballot_sync(0xffffffffu, valid[3]);
            {
                int idx = woff + __popc(b0 & lt_mask) + __popc(b1 & lt_mask) + __popc(b2 & lt_mask) + __popc(b3 & lt_mask);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (valid[i]) {
                        nu_c[idx] = make_float2(n1[i], n2[i]);
                        dp_c[idx] = make_uint2(__float_as_uint(dd[i]), pk0 + (uint32_t)i);
                    }
                    idx += valid[i] ? 1 : 0;
                }
            }
            woff += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
            if (FAST) {
                // every lane has consumed its part of the stage: hand it back to the copy engine
                __syncwarp();
                if (cpa) issue_cpa(k + kRingStages);
                else if (k + kRingStages < my_cnt) issue(k + kRingStages);
            }
        }
        if (FAST && cpa) asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (c.enable) {
            if (lane == 0) a.sc.chunk_cnt[(long long)slot * kVtMaxChunks + gchunk] = woff;
            if (a.wl_pixels) {
                const int cs = warp_sum(cand_cnt);
                if (lane == 0 && cs) atomicAdd(a.wl_pixels + t, cs);
            }
        }
    }
    if (!c.enable) {  // propagation only (uniform over the cluster)
        if (rank == 0 && warp == 0) {
            if (lane == 0 && a.wl_units) a.wl_units[t] = n_list;
            check_out();
        }
        return;
    }
    cluster_barrier();      // ---- #2: records and chunk counts written
    if (clk) clk[2] = global_timer();

    // valid-rank base of every chunk
    for (int i = tid; i < n_chunks; i += kVtThreads) s_cnt[i] = __ldcg(a.sc.chunk_cnt + (long long)slot * kVtMaxChunks + i);
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < n_chunks; ++i) {
            s_base[i] = acc;
            acc += s_cnt[i];
        }
        s_base[n_chunks] = acc;
    }
    __syncthreads();
    const int N = s_base[n_chunks];
    SelParams sp;
    sp.m = 0.f; sp.k2 = 0.f; sp.floor_ = 0.f; sp.use = 0; sp.b = 0.0;

    if (a.weight_flow && N > 0) {
        // ============================ pair + first radix level ==================================
        // r_i = sqrt(nu[i]^2 + nu[N+i]^2) over the interleaved innovation vector (Q3); one lane item = the two r of an
        // even i: own record i/2 and the partner record(s) holding elements N+i, N+i+1.
        for (int i = tid; i < kSelBins; i += kVtThreads) s_hist[i] = 0;
        __syncthreads();
        const uint32_t h_s = smem_u32(s_hist);
        {
            const int items = (N + 1) >> 1;
            constexpr int PU = 4, PB = 32 * PU;  // PU trips of 32 items in flight per warp: all loads issued before the first use
            const int per_w = ((items + n_chunks - 1) / n_chunks + PB - 1) & ~(PB - 1);
            const int j0 = gchunk * per_w, j1 = min(items, j0 + per_w);
            const bool odd = N & 1;
            const int qoff = N >> 1;  // even N: partner record of item j is qoff + j; odd N: elements straddle qoff + j, qoff + j + 1
            int c_own = 0, c_par = 0;  // (warp-uniform) chunks holding the first own / partner record of the block
            if (j0 < j1) {
                while (j0 >= s_base[c_own + 1]) ++c_own;
                while (qoff + j0 >= s_base[c_par + 1]) ++c_par;
            }
            for (int jb = j0; jb < j1; jb += PB) {
                const int je = min(jb + PB, j1);
                float2 own[PU], q0[PU];
                float q1x[PU];
                // records inside one chunk are contiguous: a block that does not cross a chunk boundary (almost all of
                // them) is addressed from two warp-uniform pointers; the others locate every record by search
                const int p_last = qoff + je - 1 + ((odd && 2 * (je - 1) + 1 < N) ? 1 : 0);
                if (je - 1 < s_base[c_own + 1] && p_last < s_base[c_par + 1]) {
                    const float2* own_p = nu_t + ((long long)c_own * chunk_stride - s_base[c_own]);
                    const float2* par_p = nu_t + ((long long)c_par * chunk_stride - s_base[c_par] + qoff);
#pragma unroll
                    for (int u = 0; u < PU; ++u) {
                        const int j = jb + 32 * u + lane;
                        own[u] = make_float2(0.f, 0.f);
                        q0[u] = own[u];
                        q1x[u] = 0.f;
                        if (j < je) {
                            own[u] = __ldcg(own_p + j);
                            q0[u] = __ldcg(par_p + j);
                            if (odd && 2 * j + 1 < N) q1x[u] = __ldcg(reinterpret_cast<const float*>(par_p + j + 1));
                        }
                    }
                } else {
                    int co = c_own, cp = c_par;  // per-lane hints: the items of a lane are visited in ascending order
#pragma unroll
                    for (int u = 0; u < PU; ++u) {
                        const int j = jb + 32 * u + lane;
                        own[u] = make_float2(0.f, 0.f);
                        q0[u] = own[u];
                        q1x[u] = 0.f;
                        if (j < je) {
                            own[u] = __ldcg(nu_t + rec_index(j, co, s_base, chunk_stride));
                            q0[u] = __ldcg(nu_t + rec_index(qoff + j, cp, s_base, chunk_stride));
                            if (odd && 2 * j + 1 < N) {
                                int cp2 = cp;
                                q1x[u] = __ldcg(nu_t + rec_index(qoff + j + 1, cp2, s_base, chunk_stride)).x;
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < PU; ++u) {
                    const int j = jb + 32 * u + lane;
                    if (j < je) {
                        const int i0 = 2 * j;
                        // even N: elements N+i0, N+i0+1 are both components of record N/2 + j; odd N: element N+i0 is
                        // the second component of record (N-1)/2 + j, element N+i0+1 the first of the next record
                        const float pa = odd ? q0[u].y : q0[u].x;
                        const float pb = odd ? q1x[u] : q0[u].y;
                        const float ra = sqrt_approx(fmaf(own[u].x, own[u].x, pa * pa));
                        const float rb = sqrt_approx(fmaf(own[u].y, own[u].y, pb * pb));
                        if (i0 + 1 < N) {
                            *reinterpret_cast<float2*>(r_t + i0) = make_float2(ra, rb);
                            red_shared_inc(h_s + (((__float_as_uint(rb) >> 19) & 0xfffu) << 2));
                        } else {
                            r_t[i0] = ra;
                        }
                        red_shared_inc(h_s + (((__float_as_uint(ra) >> 19) & 0xfffu) << 2));
                    }
                }
                if (je < j1) {  // chunks of the next block's first item
                    while (je >= s_base[c_own + 1]) ++c_own;
                    while (qoff + je >= s_base[c_par + 1]) ++c_par;
                }
            }
        }
        __syncthreads();
        uint32_t* gh = a.sc.hist + (long long)slot * 3 * kSelBins;
        for (int i = tid; i < kSelBins; i += kVtThreads) {
            const uint32_t v = s_hist[i];
            if (v) atomicAdd(gh + i, v);
        }
        cluster_barrier();  // ---- #3: r written, level-0 histogram complete
        if (clk) clk[3] = global_timer();

        // bin scan by the whole block: the bin holding rank k among `nb` bins (16 per thread)
        auto scan_bins = [&](const uint32_t* hsrc, uint32_t k, uint32_t& bin_out, uint32_t& k_out) {
            constexpr int PER = kSelBins / kVtThreads;
            uint32_t loc[PER];
            int sum = 0;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                loc[i] = __ldcg(hsrc + tid * PER + i);
                sum += (int)loc[i];
            }
            int total;
            uint32_t before = (uint32_t)block_excl_scan(sum, s_w, total);
            if (k >= before && k < before + (uint32_t)sum) {  // exactly one thread
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    if (k < before + loc[i]) {
                        s_misc[1] = tid * PER + i;
                        s_misc[2] = (int)(k - before);
                        break;
                    }
                    before += loc[i];
                }
            }
            __syncthreads();
            bin_out = (uint32_t)s_misc[1];
            k_out = (uint32_t)s_misc[2];
            __syncthreads();
        };
        // this CTA's share of the r array (whole float4s; the tail goes to the last CTA)
        const int n4 = N >> 2;
        const int per4 = (n4 + (int)C - 1) / (int)C;
        const int lo4 = min(n4, (int)rank * per4), hi4 = min(n4, lo4 + per4);
        const uint4* keys4 = reinterpret_cast<const uint4*>(r_t);
        const uint32_t* keys = reinterpret_cast<const uint32_t*>(r_t);
        const bool has_tail = rank == C - 1;

        // ============================ second radix level =========================================
        uint32_t bin0, k0r;
        scan_bins(gh, (uint32_t)(N >> 1), bin0, k0r);  // upper median = s[N/2]
        for (int i = tid; i < kSelBins; i += kVtThreads) s_hist[i] = 0;
        __syncthreads();
        {
            auto count1 = [&](uint32_t key) { red_shared_inc_if_eq(h_s + (((key >> 7) & 0xfffu) << 2), key >> 19, bin0); };
            int i4 = lo4 + tid;
            for (; i4 + kVtThreads < hi4; i4 += 2 * kVtThreads) {
                const uint4 ka = __ldcg(keys4 + i4), kb = __ldcg(keys4 + i4 + kVtThreads);
                count1(ka.x); count1(ka.y); count1(ka.z); count1(ka.w);
                count1(kb.x); count1(kb.y); count1(kb.z); count1(kb.w);
            }
            if (i4 < hi4) {
                const uint4 ka = __ldcg(keys4 + i4);
                count1(ka.x); count1(ka.y); count1(ka.z); count1(ka.w);
            }
            if (has_tail)
                for (int i = (n4 << 2) + tid; i < N; i += kVtThreads) count1(__ldcg(keys + i));
        }
        __syncthreads();
        for (int i = tid; i < kSelBins; i += kVtThreads) {
            const uint32_t v = s_hist[i];
            if (v) atomicAdd(gh + kSelBins + i, v);
        }
        cluster_barrier();  // ---- #4
        if (clk) clk[4] = global_timer();

        // ============================ last level + statistics ====================================
        uint32_t bin1, k1r;
        scan_bins(gh + kSelBins, k0r, bin1, k1r);
        const uint32_t pbin = (bin0 << 12) | bin1;  // == key >> 7 of the upper median
        if (tid < 128) s_hist[tid] = 0;
        __syncthreads();
        {
            float tot = 0.f, ls = 0.f, lm = 0.f;  // per-thread FP32 partials (a few hundred terms), FP64 across threads
            unsigned lc = 0;
            auto visit = [&](uint32_t key) {
                const float v = __uint_as_float(key);
                const uint32_t kb = key >> 7;
                const bool below = kb < pbin;
                tot += v;
                ls += below ? v : 0.f;
                lc += below ? 1u : 0u;
                lm = fmaxf(lm, below ? v : 0.f);
                red_shared_inc_if_eq(h_s + ((key & 0x7fu) << 2), kb, pbin);
            };
            int i4 = lo4 + tid;
            for (; i4 + kVtThreads < hi4; i4 += 2 * kVtThreads) {
                const uint4 ka = __ldcg(keys4 + i4), kb = __ldcg(keys4 + i4 + kVtThreads);
                visit(ka.x); visit(ka.y); visit(ka.z); visit(ka.w);
                visit(kb.x); visit(kb.y); visit(kb.z); visit(kb.w);
            }
            if (i4 < hi4) {
                const uint4 ka = __ldcg(keys4 + i4);
                visit(ka.x); visit(ka.y); visit(ka.z); visit(ka.w);
            }
            if (has_tail)
                for (int i = (n4 << 2) + tid; i < N; i += kVtThreads) visit(__ldcg(keys + i));
            double dtot = warp_sum((double)tot), dls = warp_sum((double)ls);
            lc = (unsigned)warp_sum((int)lc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lm = fmaxf(lm, __shfl_xor_sync(0xffffffffu, lm, o));
            if (lane == 0) {
                s_part[warp][0] = dtot;
                s_part[warp][1] = dls;
                s_part[warp][2] = (double)lc;
                s_part[warp][3] = (double)lm;
            }
            __syncthreads();
            if (tid == 0) {  // fixed order: deterministic
                double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
                for (int w = 0; w < kVtWarps; ++w) {
                    p0 += s_part[w][0];
                    p1 += s_part[w][1];
                    p2 += s_part[w][2];
                    p3 = fmax(p3, s_part[w][3]);
                }
                double* o = a.sc.sel_part + ((long long)slot * kVtMaxCluster + rank) * 4;
                o[0] = p0; o[1] = p1; o[2] = p2; o[3] = p3;
            }
            if (tid < 128) {
                const uint32_t v = s_hist[tid];
                if (v) atomicAdd(gh + 2 * kSelBins + tid, v);
            }
        }
        cluster_barrier();  // ---- #5
        if (clk) clk[5] = global_timer();
        // every CTA is past its scans of the level-0 / level-1 histograms: back to zero for the slot's next user
        for (int i = (int)rank * kVtThreads + tid; i < 2 * kSelBins; i += (int)C * kVtThreads) gh[i] = 0;

        // ============================ median, b (every CTA, warp 0) ==============================
        if (warp == 0) {
            double total_sum = 0.0, less_sum = 0.0, less_cnt = 0.0, less_max = 0.0;
            for (unsigned r = 0; r < C; ++r) {
                const double* o = a.sc.sel_part + ((long long)slot * kVtMaxCluster + r) * 4;
                total_sum += __ldcg(o);
                less_sum += __ldcg(o + 1);
                less_cnt += __ldcg(o + 2);
                less_max = fmax(less_max, __ldcg(o + 3));
            }
            // 128 bins of the low 7 key bits inside the 24-bit prefix: 4 per lane; keys sharing all 31 bits are the same float
            uint32_t loc[4];
            uint32_t sum = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                loc[i] = __ldcg(gh + 2 * kSelBins + lane * 4 + i);
                sum += loc[i];
            }
            const uint32_t incl = (uint32_t)warp_scan_incl((int)sum, lane);
            const uint32_t before = incl - sum;
            int bsel = -1;
            if (k1r >= before && k1r < before + sum) {
                uint32_t acc = before;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (bsel < 0 && k1r < acc + loc[i]) bsel = lane * 4 + i;
                    acc += loc[i];
                }
            }
            const uint32_t vote = __ballot_sync(0xffffffffu, bsel >= 0);
            const int src = __ffs(vote) - 1;
            bsel = __shfl_sync(0xffffffffu, bsel, src < 0 ? 0 : src);
            double c_in = 0.0, s_in = 0.0;
            float m_in = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = lane * 4 + i;
                if (b < bsel && loc[i]) {
                    const float v = __uint_as_float((pbin << 7) | (uint32_t)b);
                    c_in += (double)loc[i];
                    s_in += (double)loc[i] * (double)v;
                    m_in = fmaxf(m_in, v);
                }
            }
            c_in = warp_sum(c_in);
            s_in = warp_sum(s_in);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m_in = fmaxf(m_in, __shfl_xor_sync(0xffffffffu, m_in, o));
            less_cnt += c_in;
            less_sum += s_in;
            less_max = fmax(less_max, (double)m_in);
            // SKFCorrection.cpp:95-102: median (even: mean of the two middle values), b = mean |r - m|
            const double n = (double)N;
            const double kh = (double)(N >> 1);
            const double v1 = (double)__uint_as_float((pbin << 7) | (uint32_t)bsel);
            const double lower = (less_cnt == kh) ? less_max : v1;
            const bool even = (N & 1) == 0;
            const double m = even ? 0.5 * (lower + v1) : v1;
            const double s_below = less_sum + (kh - less_cnt) * v1;  // sum of the N/2 smallest
            const double s_above = total_sum - s_below;
            const double b = ((s_above - (n - kh) * m) + (kh * m - s_below)) / n;
            if (lane == 0) {
                SelParams o;
                o.m = (float)m;
                o.b = b;
                o.use = b > 1e-4 ? 1 : 0;  // SKFCorrection.cpp:106
                o.k2 = o.use ? (float)(-1.4426950408889634 / b) : 0.f;
                o.floor_ = (float)(2e-6 * b);
                s_sp = o;
            }
        }
        __syncthreads();
        sp = s_sp;
    }

    // ================================ pass B =======================================================
    // The N records are split evenly over the cluster's warps by RANK (the chunks of pass A hold unequal numbers of
    // valid pixels); a warp walks the one to three chunk segments its rank range covers.  The 40 sums use the
    // UN-normalised weights l' = max(exp(-|n - m| / b), 2e-6 b); the true maximum of the likelihoods
    // (SKFCorrection.cpp:114) follows from the smallest |n - m| seen and is divided out in the epilogue.
    {
        const int per_b = (N + n_chunks - 1) / n_chunks;
        const int lo0 = min(N, gchunk * per_b), hi0 = min(N, lo0 + per_b);
        int c0 = 0;  // chunk holding rank lo0
        for (int i = lane; i < n_chunks; i += 32) c0 += (s_base[i + 1] <= lo0) ? 1 : 0;
        c0 = min(warp_sum(c0), n_chunks - 1);
        float dmin = 3.0e38f;
        const bool fp64 = a.accum_fp64 == 1 || (a.accum_fp64 == 2 && N < kAutoFp64Candidates);
        double* out = &s_part[warp][0];
        if (!fp64) {
            // SW = -1: all 40 sums in one sweep over the records.  SW = 0 / 1 (the 80-register build, three CTAs per SM):
            // two sweeps with 20 sums each - row 0 of the S blocks and the g vectors, then rows 1-4 - so that the accumulators
            // fit the register budget; the records are re-read from L2 and the weight is recomputed
            auto run_sweep = [&](auto SWC) {
            constexpr int SW = decltype(SWC)::value;
            constexpr int PF = SW < 0 ? 4 : 2;  // records in flight per lane
            float2 acc2[20];
#pragma unroll
            for (int i = 0; i < 20; ++i) acc2[i] = make_float2(0.f, 0.f);
            auto accum = [&](float n1, float n2, uint32_t dbits, uint32_t pk, bool ok) {
                pk = ok ? pk : 0u;
                const float xh = s_xh[pk & 0xffffu], yh = s_yh[pk >> 16];
                const float d = ok ? __uint_as_float(dbits) : 1.0f;
                n1 = ok ? n1 : 0.f;
                n2 = ok ? n2 : 0.f;
                const float ia = rcp_approx(d);
                const float nr = sqrt_approx(fmaf(n1, n1, n2 * n2));
                const float dev = fabsf(nr - sp.m);
                if (SW <= 0) dmin = fminf(dmin, ok ? dev : 3.0e38f);
                float l = sp.use ? fmaxf(ex2_approx(dev * sp.k2), sp.floor_) : 1.0f;
                l = ok ? l : 0.f;
                const float2 e[5] = {make_float2(ia, ia), make_float2(-xh * ia, -yh * ia), make_float2(-xh * yh, -fmaf(yh, yh, 1.0f)),
                                     make_float2(fmaf(xh, xh, 1.0f), xh * yh), make_float2(-yh, xh)};
                const float2 ll = make_float2(l, l);
                float2 w[5];
#pragma unroll
                for (int kk = 0; kk < 5; ++kk) w[kk] = __fmul2_rn(ll, e[kk]);
                int o = 0;
#pragma unroll
                for (int r = 0; r < 5; ++r)
#pragma unroll
                    for (int qq = r; qq < 5; ++qq) {
                        if (SW < 0 || (SW == 0) == (r == 0)) acc2[o] = __ffma2_rn(w[r], e[qq], acc2[o]);
                        ++o;
                    }
                if (SW <= 0) {
                    const float2 zz = make_float2(n1, n2);
#pragma unroll
                    for (int kk = 0; kk < 5; ++kk) acc2[15 + kk] = __ffma2_rn(w[kk], zz, acc2[15 + kk]);
                }
            };
            // software pipeline: the records of the NEXT trip are in flight while this trip's are accumulated
            auto fetch = [&](const float2* nup, const uint2* dpp, int k, int n, float2 (&A)[PF], uint2 (&D)[PF]) {
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    A[u] = make_float2(0.f, 0.f);
                    D[u] = make_uint2(0u, 0u);
                    if (k + 32 * u < n) {
                        A[u] = __ldcg(nup + k + 32 * u);
                        D[u] = __ldcg(dpp + k + 32 * u);
                    }
                }
            };
            int lo = lo0, c = c0;
#pragma unroll 1
            while (lo < hi0) {
                const int ce = min(hi0, s_base[c + 1]);
                const int n = ce - lo;
                const long long off = (long long)c * chunk_stride + (lo - s_base[c]);
                const float2* nup = nu_t + off;
                const uint2* dpp = dp_t + off;
                // two register sets alternate (no copies): while one trip is accumulated the next is in flight
                float2 A[PF], B[PF];
                uint2 D[PF], E[PF];
                fetch(nup, dpp, lane, n, A, D);
#pragma unroll 1
                for (int k = lane; k < n; k += 64 * PF) {
                    fetch(nup, dpp, k + 32 * PF, n, B, E);
#pragma unroll
                    for (int u = 0; u < PF; ++u) accum(A[u].x, A[u].y, D[u].x, D[u].y, k + 32 * u < n);
                    fetch(nup, dpp, k + 64 * PF, n, A, D);
#pragma unroll
                    for (int u = 0; u < PF; ++u) accum(B[u].x, B[u].y, E[u].x, E[u].y, k + 32 * PF + 32 * u < n);
                }
                lo = ce;
                ++c;
            }
#pragma unroll
            for (int o = 0; o < 15; ++o) {
                if (!(SW < 0 || (SW == 0) == (o < 5))) continue;  // (o < 5 <=> row 0)
                const float s1 = warp_sum(acc2[o].x), s2 = warp_sum(acc2[o].y);
                if (lane == 0) {
                    out[o] = (double)s1;
                    out[15 + o] = (double)s2;
                }
            }
            if (SW <= 0) {
#pragma unroll
                for (int kk = 0; kk < 5; ++kk) {
                    const float s1 = warp_sum(acc2[15 + kk].x), s2 = warp_sum(acc2[15 + kk].y);
                    if (lane == 0) {
                        out[30 + kk] = (double)s1;
                        out[35 + kk] = (double)s2;
                    }
                }
            }
            };
            if (REGS > 80) {
                run_sweep(std::integral_constant<int, -1>());
            } else {
                run_sweep(std::integral_constant<int, 0>());
                run_sweep(std::integral_constant<int, 1>());
            }
        } else {
            // FP64 terms and sums (small, typically ill-conditioned tracks; accum_fp64): three sweeps over the chunk's
            // records with 15 / 15 / 10 accumulators keep the register count of the common FP32 path
            const double ifx = a.inv_fx, ify = a.inv_fy;
#pragma unroll 1
            for (int sweep = 0; sweep < 3; ++sweep) {
                double acc[15];
#pragma unroll
                for (int i = 0; i < 15; ++i) acc[i] = 0.0;
                int cc = c0;
#pragma unroll 1
                for (int k = lo0 + lane; k < hi0; k += 32) {
                    const long long ri = rec_index(k, cc, s_base, chunk_stride);
                    const float2 nn = __ldcg(nu_t + ri);
                    const uint2 dpk = __ldcg(dp_t + ri);
                    const double xh = ((double)(dpk.y & 0xffffu) - a.cx) * ifx, yh = ((double)(dpk.y >> 16) - a.cy) * ify;
                    const float df = __uint_as_float(dpk.x);
                    // 1/d from the FP32 approximation (rel. error < 2^-22) by two Newton steps
                    const double dd = (double)df;
                    double r = (double)rcp_approx(df);
                    r = fma(r, fma(-dd, r, 1.0), r);
                    r = fma(r, fma(-dd, r, 1.0), r);
                    const float nr = sqrt_approx(fmaf(nn.x, nn.x, nn.y * nn.y));
                    const float dev = fabsf(nr - sp.m);
                    dmin = fminf(dmin, dev);
                    const double l = sp.use ? (double)fmaxf(ex2_approx(dev * sp.k2), sp.floor_) : 1.0;
                    double e[5];
                    if (sweep == 0) {
                        e[0] = r; e[1] = -xh * r; e[2] = -xh * yh; e[3] = 1.0 + xh * xh; e[4] = -yh;
                    } else {
                        e[0] = r; e[1] = -yh * r; e[2] = -(1.0 + yh * yh); e[3] = xh * yh; e[4] = xh;
                    }
                    if (sweep < 2) {
                        int o = 0;
#pragma unroll
                        for (int rr = 0; rr < 5; ++rr) {
                            const double wl = l * e[rr];
#pragma unroll
                            for (int qq = rr; qq < 5; ++qq) {
                                acc[o] = fma(wl, e[qq], acc[o]);
                                ++o;
                            }
                        }
                    } else {
                        const double e1[5] = {r, -xh * r, -xh * yh, 1.0 + xh * xh, -yh};
#pragma unroll
                        for (int kk = 0; kk < 5; ++kk) {
                            acc[kk] = fma(l * e1[kk], (double)nn.x, acc[kk]);
                            acc[5 + kk] = fma(l * e[kk], (double)nn.y, acc[5 + kk]);
                        }
                    }
                }
                const int nacc = sweep < 2 ? 15 : 10;
#pragma unroll
                for (int i = 0; i < 15; ++i) {
                    const double s = warp_sum(acc[i]);
                    if (lane == 0 && i < nacc) out[sweep * 15 + i] = s;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        if (lane == 0) {
            out[40] = (double)(hi0 - lo0);
            out[41] = (double)dmin;
        }
        __syncthreads();
        if (tid < kVtPartN) {  // fixed order over the warps: deterministic
            double s = tid == 41 ? 3.0e38 : 0.0;
            for (int w = 0; w < kVtWarps; ++w) s = tid == 41 ? fmin(s, s_part[w][41]) : s + s_part[w][tid];
            a.sc.part[((long long)t * kVtMaxCluster + rank) * kVtPartN + tid] = s;
        }
    }
    cluster_barrier();  // ---- #6: partials written; the scratch records are dead from here on
    if (clk) clk[6] = global_timer();
    if (rank != 0) return;

    // the 6x6 solve is a separate one-warp-per-track kernel (k_velocity_epilogue): done here it would hold the
    // cluster's slot for a serial FP64 tail ten times longer than the partial sums it needs
    if (warp != 0) return;
    if (lane == 0) {
        double* ts = a.sc.track_sel + (long long)t * 4;
        ts[0] = sp.use ? 1.0 : 0.0;
        ts[1] = sp.b;
        ts[2] = (double)N;
        ts[3] = (double)C;
        if (a.weight_flow && N > 0) {  // level-2 histogram back to zero for the slot's next user (levels 0, 1: see above)
            uint32_t* gh2 = a.sc.hist + (long long)slot * 3 * kSelBins + 2 * kSelBins;
            for (int i = 0; i < 128; ++i) gh2[i] = 0;
        }
        __threadfence();
        // release the scratch slot: every CTA of the cluster is past its last access
        atomicAnd(a.sc.slot_bitmap + (slot >> 5), ~(1u << (slot & 31)));
    }
    __syncwarp();
    if (clk) clk[7] = global_timer();
    if (lane == 0) span_stamp(a.span_clock, true);
    check_out();
}

// ---- per-track epilogue: FP64 sum of the CTA partials, 6x6 solve, observability gate, publish ----------------
// One warp per track.  Tracks that took no part this step publish their (unchanged) twist and a zero count.
struct EpiArgs {
    int n_tracks;
    const VelCtl* ctl;
    const double* part; const double* track_sel;
    const double* x_pred;
    double* v_mean; double* v_cov; const double* q_diag;
    double r0, r1, fx, fy;
    double* vel_hist; int hist_ring;
    int32_t* out_count; double* out_lambda; double* out_eta;
    int update_state;
};
constexpr int kEpiWarps = 4;

__global__ void __launch_bounds__(32 * kEpiWarps) k_velocity_epilogue(const EpiArgs a) {
    __shared__ double s_sums_all[kEpiWarps][kVtPartN], sLm_all[kEpiWarps][36], sEta_all[kEpiWarps][6], sRhs_all[kEpiWarps][6];
    __shared__ double M1_all[kEpiWarps][6][12], M2_all[kEpiWarps][6][12];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kEpiWarps + w;
    if (t >= a.n_tracks) return;
    double* s_sums = s_sums_all[w];
    double* sLm = sLm_all[w];
    double* sEta = sEta_all[w];
    double* sRhs = sRhs_all[w];
    double (*M1)[12] = M1_all[w];
    double (*M2)[12] = M2_all[w];
    const VelCtl c = a.ctl[t];
    double* xs = a.v_mean + (long long)t * 6;
    double* P = a.v_cov + (long long)t * 36;
    int count = 0;
    if (c.enable) {
        const double* ts = a.track_sel + (long long)t * 4;
        const int use = ts[0] != 0.0;
        const double b = ts[1];
        count = (int)ts[2];
        const int C = (int)ts[3];
        for (int i = lane; i < kVtPartN; i += 32) {  // fixed order over the CTAs: deterministic
            double s = i == 41 ? 3.0e38 : 0.0;
            for (int r = 0; r < C; ++r) {
                const double v = a.part[((long long)t * kVtMaxCluster + r) * kVtPartN + i];
                s = i == 41 ? fmin(s, v) : s + v;
            }
            s_sums[i] = s;
        }
        __syncwarp();
        // likelihoods divided by their maximum (SKFCorrection.cpp:114): l_j / l_max = l'_j / max(exp(-dmin / b), 2e-6 b)
        double scale = 1.0;
        if (use) scale = 1.0 / fmax(exp(-s_sums[41] / b), 2e-6 * b);
        const double k1 = scale * (a.fx * c.dt) * (a.fx * c.dt) / a.r0, k2 = scale * (a.fy * c.dt) * (a.fy * c.dt) / a.r1;
        const double e1 = scale * (a.fx * c.dt) / a.r0, e2 = scale * (a.fy * c.dt) / a.r1;
        // Lambda(r, c): S1 lives on the index set {0,2,3,4,5}, S2 on {1,2,3,4,5}; entry (i, j), i <= j, of the packed
        // upper triangle of a 5x5 is i*5 - i*(i-1)/2 + (j - i)
        for (int e = lane; e < 36; e += 32) {
            const int r = e / 6, cc = e - r * 6;
            double v = 0.0;
            if (r != 1 && cc != 1) {
                const int i = r == 0 ? 0 : r - 1, j = cc == 0 ? 0 : cc - 1;
                const int lo = min(i, j), hi = max(i, j);
                v += k1 * s_sums[lo * 5 - lo * (lo - 1) / 2 + (hi - lo)];
            }
            if (r != 0 && cc != 0) {
                const int i = r - 1, j = cc - 1;
                const int lo = min(i, j), hi = max(i, j);
                v += k2 * s_sums[15 + lo * 5 - lo * (lo - 1) / 2 + (hi - lo)];
            }
            sLm[e] = v;
        }
        if (lane < 6) {
            double v = 0.0;
            if (lane != 1) v += e1 * s_sums[30 + (lane == 0 ? 0 : lane - 1)];
            if (lane != 0) v += e2 * s_sums[35 + lane - 1];
            sRhs[lane] = v;
        }
        __syncwarp();
        // sums were taken against the innovations: eta (w.r.t. z) = sum l H^T R^-1 nu + Lambda x_pred
        if (lane < 6) {
            const double* xp = a.x_pred + (long long)t * 6;
            double v = sRhs[lane];
            for (int j = 0; j < 6; ++j) v += sLm[lane * 6 + j] * xp[j];
            sEta[lane] = v;
        }
        __syncwarp();
    } else {
        for (int e = lane; e < 36; e += 32) sLm[e] = 0.0;
        if (lane < 6) sEta[lane] = 0.0;
        __syncwarp();
    }
    if (a.out_lambda)
        for (int i = lane; i < 36; i += 32) a.out_lambda[(long long)t * 36 + i] = sLm[i];
    if (a.out_eta && lane < 6) a.out_eta[(long long)t * 6 + lane] = sEta[lane];
    // ROFTFilter.cpp:294-301: fewer than 3 valid pixels (or an empty measurement, SKFCorrection.cpp:60-68 keeps the
    // PREDICTED state, which the observability gate then reverts) -> the belief is left untouched.
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
            for (int j = 0; j < 6; ++j) v += M1[lane][6 + j] * xs[j];
            sRhs[lane] = v;
        }
        __syncwarp();
        if (lane < 6) {
            double v = 0.0;
            for (int j = 0; j < 6; ++j) v += M2[lane][6 + j] * sRhs[j];
            xs[lane] = v;
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
        h[lane] = xs[lane];
    }
}

}  // namespace

// shared-memory layout of the cluster kernel for a geometry / cluster size
static void vt_smem_layout(const Geom& g, int n_units, int cluster, VtArgs& va, int& bytes) {
    int off = kRingRegion;
    off = (off + 15) & ~15;
    va.smem_tab = off;
    off += (g.W + g.H) * 4;
    off = (off + 15) & ~15;
    const int n_chunks = cluster * kVtWarps;
    const int per_chunk_max = (n_units + n_chunks - 1) / n_chunks;
    va.slice_cap = kVtWarps * per_chunk_max;
    va.smem_list = off;
    off += va.slice_cap * 4;
    bytes = off;
}

int velocity_cluster_size() {
    static const int v = [] {
        const char* e = getenv("ROFTB_CLUSTER");
        int c = e ? atoi(e) : 4;  // measured best on B200 at 256 tracks: 71 clusters resident
        if (c < 1) c = 1;
        if (c > kVtMaxCluster) c = kVtMaxCluster;
        return c;
    }();
    return v;
}

// register cap the kernel is built with (see k_velocity_track)
static int velocity_regs() {
    static const int v = [] {
        const char* e = getenv("ROFTB_VT_REGS");
        const int r = e ? atoi(e) : 128;
        return r <= 80 ? 80 : r >= 128 ? 128 : r >= 112 ? 112 : r >= 104 ? 104 : 96;
    }();
    return v;
}

using VtKernel = void (*)(const VtArgs);
static VtKernel vt_kernel(bool fast) {
    switch (velocity_regs()) {
        case 80: return fast ? k_velocity_track<true, 80> : k_velocity_track<false, 80>;
        case 128: return fast ? k_velocity_track<true, 128> : k_velocity_track<false, 128>;
        case 112: return fast ? k_velocity_track<true, 112> : k_velocity_track<false, 112>;
        case 104: return fast ? k_velocity_track<true, 104> : k_velocity_track<false, 104>;
        default: return fast ? k_velocity_track<true, 96> : k_velocity_track<false, 96>;
    }
}

int velocity_prepare_device(const Geom& g, int n_units, int* max_active_clusters) {
    const int cluster = velocity_cluster_size();
    VtArgs va;
    int bytes = 0;
    vt_smem_layout(g, n_units, cluster, va, bytes);
    for (int fast = 0; fast < 2; ++fast) {
        if (cudaFuncSetAttribute(vt_kernel(fast != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return -1;
        cudaFuncSetAttribute(vt_kernel(fast != 0), cudaFuncAttributeNonPortableClusterSizeAllowed, 1);  // (4x the base cluster may be 16)
    }
    if (max_active_clusters) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(cluster, 1024);
        cfg.blockDim = dim3(kVtThreads);
        cfg.dynamicSmemBytes = bytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cluster;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, vt_kernel(true), &cfg);
        if (e != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = 148 * (velocity_regs() <= 80 ? 3 : 2) / cluster + 1;
        }
        *max_active_clusters = n;
    }
    return 0;
}

int launch_velocity(const VelocityArgs& a, cudaStream_t s) {
    const int T = a.n_tracks;
    const Geom& g = a.g;
    const int n_units = (g.HW + kUnitPx - 1) / kUnitPx;
    const int cluster = velocity_cluster_size();
    VtArgs va;
    memset(&va, 0, sizeof(va));
    va.g = g;
    va.ft = a.ft;
    va.seg = a.seg;
    va.seg_stride = a.seg_stride;
    va.thr = a.thr;
    va.occ_src = a.occ_src;
    va.ctl = a.ctl;
    va.n_units = n_units;
    va.weight_flow = a.weight_flow;
    va.accum_fp64 = a.accum_fp64;
    va.update_state = a.update_state;
    va.x_pred = a.x_pred_override ? a.x_pred_override : a.v_mean;
    va.fx = a.fx; va.fy = a.fy; va.cx = a.cx; va.cy = a.cy;
    va.inv_fx = 1.0 / a.fx; va.inv_fy = 1.0 / a.fy;
    va.plan = a.fuse_scatter ? a.plan : nullptr;
    va.state_dst = a.state_dst;
    va.occ_dst = a.occ_dst;
    va.sc = a.scratch;
    va.wl_units = a.wl_units; va.wl_pixels = a.wl_pixels;
    va.order = a.order; va.order_next = a.order_next;
    va.done_ticket = (a.order_next && a.wl_units) ? a.done_ticket : nullptr;
    va.phase_clock = a.phase_clock;
    va.span_clock = a.span_clock;
    static const int env_hint = [] { const char* e = getenv("ROFTB_L2HINT"); return e ? atoi(e) : 1; }();
    va.l2_hint = env_hint;
    static const int env_stage = [] { const char* e = getenv("ROFTB_STAGE"); return e ? atoi(e) : 1; }();
    va.stage_mode = env_stage;
    // the bulk-copy ring needs dense float2 flow at full resolution and whole 16-byte groups; every other
    // configuration (CV_16SC2 / sub-sampled flow grids, stride > 1) takes the register path of the same kernel
    static const int env_ring = [] { const char* e = getenv("ROFTB_RING"); return e ? atoi(e) : 1; }();
    const bool fast = env_ring != 0 && !g.flow_s16 && g.grid == 1 && g.scale_mode == 0 && g.stride == 1 && (g.HW % 16) == 0;
    // Tracks are ordered largest first and their sizes differ by more than two orders of magnitude (an object can fill
    // the frame or a corner of it); a track's latency is inversely proportional to the CTAs that share it, and the launch
    // cannot end before its biggest track does.  ROFTB_VT_SPLIT (default 0 = off) gives the biggest tracks - by RANK,
    // which the host knows; the sizes never leave the device - a 2x / 4x larger cluster in concurrent launches of the
    // same kernel.  Measured on B200 (256 tracks): the biggest track's latency falls from 745 to 280-370 us, but kernels
    // with different cluster shapes pack worse on the GPCs and the step gets 12-15 % slower, so it stays off.
    struct Part { int first, count, cluster; cudaStream_t stream; };
    Part parts[3];
    int n_parts = 0;
    static const int env_split = [] { const char* e = getenv("ROFTB_VT_SPLIT"); return e ? atoi(e) : 0; }();
    const bool split = env_split != 0 && a.order && a.side_stream[0] && a.side_stream[1] && T >= 64 && cluster * 2 <= kVtMaxCluster;
    if (split && env_split == 1 && cluster * 4 <= kVtMaxCluster) {  // three ways: 1/16 at 4x, 1/8 at 2x, the rest
        const int n_big = T / 16, n_mid = T / 8;
        parts[n_parts++] = Part{0, n_big, cluster * 4, a.side_stream[0]};
        parts[n_parts++] = Part{n_big, n_mid, cluster * 2, a.side_stream[1]};
        parts[n_parts++] = Part{n_big + n_mid, T - n_big - n_mid, cluster, s};
    } else if (split) {  // two ways: the biggest 1/env_split of the tracks at 2x
        const int n_big = T / max(env_split, 2);
        parts[n_parts++] = Part{0, n_big, cluster * 2, a.side_stream[0]};
        parts[n_parts++] = Part{n_big, T - n_big, cluster, s};
    } else {
        parts[n_parts++] = Part{0, T, cluster, s};
    }
    if (n_parts > 1) {
        cudaEventRecord(a.side_fork, s);
        for (int i = 0; i + 1 < n_parts; ++i) cudaStreamWaitEvent(a.side_stream[i], a.side_fork, 0);
    }
    va.total_tracks = T;
    for (int pi = 0; pi < n_parts; ++pi) {
        const Part& pt = parts[pi];
        int bytes = 0;
        vt_smem_layout(g, n_units, pt.cluster, va, bytes);
        va.first = pt.first;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(pt.cluster, pt.count);
        cfg.blockDim = dim3(kVtThreads);
        cfg.dynamicSmemBytes = bytes;
        cfg.stream = pt.stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = pt.cluster;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, vt_kernel(fast), va);
        ++g_launch_count;
        if (e != cudaSuccess) return -1;
        if (pi + 1 < n_parts) cudaEventRecord(a.side_join[pi], pt.stream);
    }
    for (int pi = 0; pi + 1 < n_parts; ++pi) cudaStreamWaitEvent(s, a.side_join[pi], 0);  // join AFTER the last launch
    EpiArgs ea;
    ea.n_tracks = T;
    ea.ctl = a.ctl;
    ea.part = a.scratch.part;
    ea.track_sel = a.scratch.track_sel;
    ea.x_pred = va.x_pred;
    ea.v_mean = a.v_mean; ea.v_cov = a.v_cov; ea.q_diag = a.q_diag;
    ea.r0 = a.r_flow[0]; ea.r1 = a.r_flow[1]; ea.fx = a.fx; ea.fy = a.fy;
    ea.vel_hist = a.vel_hist; ea.hist_ring = a.hist_ring;
    ea.out_count = a.out_count; ea.out_lambda = a.out_lambda; ea.out_eta = a.out_eta;
    ea.update_state = a.update_state;
    ROFTB_LAUNCH(k_velocity_epilogue, (T + kEpiWarps - 1) / kEpiWarps, 32 * kEpiWarps, 0, s, ea);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace roftb
