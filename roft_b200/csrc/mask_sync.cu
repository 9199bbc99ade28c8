// Flow-aided mask synchronisation, batched over tracks (north-star part 2).
//
// Replaces ImageSegmentationOFAidedSource<T>::step_frame / map + cv::remap
// (src/roft-lib/include/ROFT/ImageSegmentationOFAidedSource.hpp:128-281) and the cv::threshold of
// ImageSegmentationMeasurement::freeze (src/roft-lib/src/ImageSegmentationMeasurement.cpp:61-65).
//
// The reference builds a float inverse map by forward-scattering every non-zero mask pixel through the
// buffered flow frames (last writer in row-major source order wins) and then gathers with cv::remap;
// unmapped destinations sample source pixel (0,0).  Here:
//   k_mask_stats   non-zero count / min / max of a newly delivered mask (emptiness + single-valuedness)
//   k_warp_plan    per-track decision of hpp:169-226 that depends on the mask CONTENT (empty new mask)
//                  and the per-track flow buffer bookkeeping - kept on the device so the host never syncs
//   k_warp_init    destination plane <- default value (or identity copy), winner plane <- -1; tracks whose propagation
//                  is fused into the velocity kernel are skipped (that kernel clears the destination lazily)
//   k_warp_scatter integer scatter: single-valued masks store the value byte directly (any writer wins
//                  the same value); mixed-valued masks resolve collisions with atomicMax(source index)
//   k_warp_gather  mixed-valued masks only: out(dst) = src(winner(dst))
// Every kernel that writes the mask state also maintains its OCCUPANCY FLAGS (one byte per 128-pixel unit, 1 = the unit
// holds a non-zero byte): the consumers build their worklists from the flags instead of re-reading the plane.
// All arithmetic on the chased position is IEEE FP32 add/div with x86 float->int truncation semantics
// (cvt_int) so the result is bit-exact against the reference + OpenCV.
#include "roftb_internal.cuh"

namespace roftb {

std::atomic<long long> g_launch_count{0};

namespace {

__device__ __forceinline__ const uint8_t* track_plane(const uint8_t* base, long long stride, int t) {
    return base + (long long)t * stride;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_mask_stats(const uint8_t* __restrict__ new_mask, long long new_stride,
                                                        const WarpCtl* __restrict__ ctl, int n16,
                                                        MaskStat* __restrict__ stat) {
    const int t = blockIdx.y;
    if (!ctl[t].has_new) return;
    const uint4* p = reinterpret_cast<const uint4*>(track_plane(new_mask, new_stride, t));
    int nnz = 0;
    uint32_t vmin = 0xffffffffu, vmax = 0u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) {
        uint4 w = ld_nc_u4(p + i);
        uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t nz = __vcmpne4(ws[k], 0u);
            nnz += __popc(nz) >> 3;
            vmin = __vminu4(vmin, ws[k] | ~nz);
            vmax = __vmaxu4(vmax, ws[k]);
        }
    }
    int mn = min(min(vmin & 0xff, (vmin >> 8) & 0xff), min((vmin >> 16) & 0xff, vmin >> 24));
    int mx = max(max(vmax & 0xff, (vmax >> 8) & 0xff), max((vmax >> 16) & 0xff, vmax >> 24));
    nnz = warp_sum(nnz);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0 && nnz > 0) {
        atomicAdd(&stat[t].nnz, nnz);
        atomicMin(&stat[t].vmin, mn);
        atomicMax(&stat[t].vmax, mx);
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void k_warp_plan(int n_tracks, const WarpCtl* __restrict__ ctl, MaskStat* __restrict__ stat,
                            WarpPlan* __restrict__ plan, FlowBuf* __restrict__ fbuf, const uint8_t* __restrict__ new_mask,
                            long long new_stride, const uint8_t* __restrict__ state_src, int HW, int segm_delay, int fuse) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    const WarpCtl c = ctl[t];
    MaskStat st = stat[t];
    stat[t] = MaskStat{0, 255, 0, 0};
    FlowBuf fb = fbuf[t];
    WarpPlan p;
    p.mode = kWarpCopyState;
    p.src_new = 0;
    p.zero_origin = 0;
    p.n_flows = 0;
    for (int i = 0; i < kMaxChain; ++i) p.flow_slot[i] = 0;
    p.uniform_val = fb.uniform_val;
    p.dflt = 0;
    p.fused = 0;
    p.pad = 0;
    if (c.reset) fb.n = 0;
    const int new_uniform = (st.nnz > 0 && st.vmin == st.vmax) ? st.vmin : 0;

    if (!c.flow_aided) {
        // plain source: the delivered mask simply replaces the state
        if (c.has_new) {
            p.mode = kWarpCopyNew;
            fb.uniform_val = new_uniform;
        }
    } else {
        bool valid_seg = c.has_new != 0;
        bool init_now = false;
        if (valid_seg && c.first_mask) {  // hpp:169-178: initialisation, not treated as a new mask
            init_now = true;
            valid_seg = false;
            fb.uniform_val = new_uniform;
        }
        if (valid_seg && st.nnz == 0) {  // hpp:186-197: uninformative mask is skipped
            valid_seg = false;
            if (segm_delay <= 0) fb.n = 0;
        }
        if (c.flow_valid) {  // hpp:200-209: flow_buffer_.push_back
            if (fb.n == kMaxFlows) {
                for (int i = 1; i < kMaxFlows; ++i) fb.slot[i - 1] = fb.slot[i];
                fb.n = kMaxFlows - 1;
            }
            fb.slot[fb.n++] = c.cur_slot;
        }
        if (valid_seg) {  // hpp:211-219: new mask through the last D buffered flows
            int start = 0;
            if (segm_delay > 0) start = max(0, fb.n - segm_delay);
            p.mode = kWarpScatter;
            p.src_new = 1;
            p.zero_origin = 0;
            p.n_flows = fb.n - start;
            for (int i = start; i < fb.n; ++i) p.flow_slot[i - start] = fb.slot[i];
            fb.n = 0;
            fb.uniform_val = new_uniform;
            p.uniform_val = new_uniform;
            p.dflt = track_plane(new_mask, new_stride, t)[0];
        } else if (c.flow_valid) {  // hpp:221-226: propagate with the current flow only
            p.mode = kWarpScatter;
            p.src_new = init_now ? 1 : 0;
            p.zero_origin = 1;
            p.n_flows = 1;
            p.flow_slot[0] = c.cur_slot;
            p.uniform_val = fb.uniform_val;
            p.dflt = 0;
            // state source + the current flow only + single-valued mask (plain byte stores, no winner plane)
            p.fused = (fuse && !init_now && fb.uniform_val != 0) ? 1 : 0;
        } else if (init_now) {
            p.mode = kWarpCopyNew;
        }
    }
    plan[t] = p;
    fbuf[t] = fb;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_warp_init(const WarpPlan* __restrict__ plan, const uint8_t* __restrict__ new_mask,
                                                       long long new_stride, const uint8_t* __restrict__ state_src,
                                                       uint8_t* __restrict__ state_dst, int32_t* __restrict__ winner,
                                                       const uint8_t* __restrict__ occ_src, uint8_t* __restrict__ occ_dst,
                                                       int HW, int n_units, unsigned long long* span) {
    const int t = blockIdx.y;
    const WarpPlan p = plan[t];
    if (p.fused) return;  // the velocity kernel clears (lazily) and propagates this track's mask
    span_stamp(span, false);
    const int n16 = HW >> 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* dst = reinterpret_cast<uint4*>(state_dst + (long long)t * HW);
    uint8_t* of = occ_dst ? occ_dst + (long long)t * n_units : nullptr;
    // one warp iteration = four units (512 px): one 128-bit access per lane, 8 lanes per unit
    const int n_blk = (n_units + 3) >> 2;
    const unsigned gmask = 0xffu << (8 * (lane >> 3));
    const bool scatter = p.mode == kWarpScatter;
    const bool general = scatter && p.uniform_val == 0;
    const uint32_t b = (uint32_t)p.dflt * 0x01010101u;
    const uint4 fill = make_uint4(b, b, b, b);
    int4* win = reinterpret_cast<int4*>(winner + (long long)t * HW);
    const int4 neg = make_int4(-1, -1, -1, -1);
    const uint4* src = reinterpret_cast<const uint4*>(
        p.mode == kWarpCopyNew ? track_plane(new_mask, new_stride, t) : state_src + (long long)t * HW);
    for (int blk = blockIdx.x * (kThreads / 32) + warp; blk < n_blk; blk += gridDim.x * (kThreads / 32)) {
        const int i = blk * 32 + lane;
        const int u = blk * 4 + (lane >> 3);
        const bool in = i < n16;
        if (scatter) {
            if (general) {  // mixed-valued mask: the gather writes the plane and its flags
                if (in) {
                    win[4 * i + 0] = neg;
                    win[4 * i + 1] = neg;
                    win[4 * i + 2] = neg;
                    win[4 * i + 3] = neg;
                }
            } else if (p.dflt == 0 && of) {
                // zero default: only the units the plane's previous content occupied need clearing (their flags say so)
                const bool occ = u < n_units && of[u] != 0;
                if (in && occ) dst[i] = fill;
                __syncwarp();
                if ((lane & 7) == 0 && occ) of[u] = 0;
            } else {
                if (in) dst[i] = fill;
                if (of && (lane & 7) == 0 && u < n_units) of[u] = p.dflt ? 1 : 0;
            }
        } else {
            const uint4 w = in ? ld_nc_u4(src + i) : make_uint4(0u, 0u, 0u, 0u);
            if (in) dst[i] = w;
            const unsigned any = __reduce_or_sync(gmask, w.x | w.y | w.z | w.w);
            if (of && (lane & 7) == 0 && u < n_units) of[u] = any ? 1 : 0;
        }
    }
    span_stamp(span, true);
}

// ---------------------------------------------------------------------------------------------
// Forward scatter of the non-zero pixels of a mask through a chain of flows (hpp:249-278).
// A warp takes a batch of kScUnits listed units, COMPACTS their non-zero pixels into shared memory (a mask fills about
// half of the pixels of its boundary units - chasing only the live ones halves the work and keeps every lane busy) and
// then chases them kScPix per lane at a time, so that many independent flow gathers are in flight per hop (a chain of
// D dependent DRAM accesses per pixel is pure latency otherwise).
// FASTF: float2 flow at full resolution with scale 1 - a hop is then a range test on the floats, the truncation read
// from the mantissa of t + 2^23 (no conversion instruction), one 64-bit gather and two IEEE adds.
constexpr int kScUnits = 4;   // units per batch
constexpr int kScPix = 8;     // pixels per lane per chase round

// high bit of every non-zero byte of x
__device__ __forceinline__ uint32_t nz4(uint32_t x) { return (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u; }

template <bool FASTF>
__global__ void __launch_bounds__(kThreads) k_warp_scatter(Geom g, FrameTable ft, const WarpPlan* __restrict__ plan,
                                                          const uint8_t* __restrict__ new_mask, long long new_stride,
                                                          const uint8_t* __restrict__ state_src,
                                                          uint8_t* __restrict__ state_dst, int32_t* __restrict__ winner,
                                                          const int32_t* __restrict__ s_list, const int32_t* __restrict__ s_n,
                                                          const int32_t* __restrict__ n_list, const int32_t* __restrict__ n_n,
                                                          int n_warp_tiles, uint8_t* __restrict__ occ_dst, unsigned long long* span) {
    const int t = blockIdx.y;
    __shared__ WarpPlan sp;
    __shared__ int32_t s_px[kThreads / 32][kScUnits * kUnitPx];
    if (threadIdx.x == 0) sp = plan[t];
    __syncthreads();
    if (sp.mode != kWarpScatter || sp.fused) return;
    span_stamp(span, false);
    const uint8_t* src = sp.src_new ? track_plane(new_mask, new_stride, t) : state_src + (long long)t * g.HW;
    uint8_t* dst = state_dst + (long long)t * g.HW;
    int32_t* win = winner + (long long)t * g.HW;
    uint8_t* of = occ_dst ? occ_dst + (long long)t * n_warp_tiles : nullptr;
    int last_flag = -1;
    const int nq = g.HW >> 2;
    const uint8_t uval = (uint8_t)sp.uniform_val;
    const bool general = (sp.uniform_val == 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int32_t* list = (sp.src_new ? n_list : s_list) + (long long)t * n_warp_tiles;
    const int n_list_items = (sp.src_new ? n_n : s_n)[t];
    int32_t* buf = s_px[warp];
    const int W = g.W;
    const unsigned uW = (unsigned)g.W;
    const float Wf = (float)g.W, Hf = (float)g.H, inv_w = 1.0f / (float)g.W;
    const bool small_hw = g.HW < (1 << 24);
    const unsigned sc_bias = 0x4b000000u * (uW + 1u);
    for (int li = (blockIdx.x * (kThreads / 32) + warp) * kScUnits; li < n_list_items; li += gridDim.x * (kThreads / 32) * kScUnits) {
        // ---- compaction: linear indices of the non-zero pixels of the batch, row-major ----
        int na = 0;  // (warp-uniform)
#pragma unroll
        for (int k = 0; k < kScUnits; ++k) {
            const bool in = li + k < n_list_items;
            const int q = in ? list[li + k] * 32 + lane : 0;
            uint32_t ts = (in && q < nq) ? nz4(ld_nc_u32(reinterpret_cast<const uint32_t*>(src) + q)) : 0u;
            if (sp.zero_origin && q == 0) ts &= ~0x80u;  // mask_(0,0) = 0 (hpp:224)
            const unsigned b0 = __ballot_sync(0xffffffffu, ts & 0x80u), b1 = __ballot_sync(0xffffffffu, ts & 0x8000u);
            const unsigned b2 = __ballot_sync(0xffffffffu, ts & 0x800000u), b3 = __ballot_sync(0xffffffffu, ts & 0x80000000u);
            int idx = na + __popc(b0 & lt_mask) + __popc(b1 & lt_mask) + __popc(b2 & lt_mask) + __popc(b3 & lt_mask);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (ts & (0x80u << (8 * i))) buf[idx++] = (q << 2) + i;
            na += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
        }
        __syncwarp();
        // ---- chase ----
        for (int base = 0; base < na; base += 32 * kScPix) {
            float tx[kScPix], ty[kScPix];
            bool alive[kScPix];
            int spx[kScPix];
#pragma unroll
            for (int i = 0; i < kScPix; ++i) {
                const int e = base + i * 32 + lane;
                alive[i] = e < na;
                const int px = alive[i] ? buf[e] : 0;
                spx[i] = px;
                int v, u;
                if (small_hw) {  // reciprocal multiply with a +-1 fix-up (exact for HW < 2^24)
                    v = (int)((float)px * inv_w);
                    u = px - v * W;
                    if (u < 0) { u += W; --v; }
                    if (u >= W) { u -= W; ++v; }
                } else {
                    v = px / W;
                    u = px - v * W;
                }
                tx[i] = (float)u;
                ty[i] = (float)v;
            }
            for (int h = 0; h < sp.n_flows; ++h) {
                const char* base_f = reinterpret_cast<const char*>(ft.flow[sp.flow_slot[h]]) +
                                     (long long)t * ft.flow_stride * (g.flow_s16 ? 2 : 4);
                float2 f[kScPix];
#pragma unroll
                for (int i = 0; i < kScPix; ++i) {
                    f[i] = make_float2(0.f, 0.f);
                    if (FASTF) {
                        // (int)t in [0, n) <=> -1 < t < n under C truncation (NaN fails; +-inf / huge fail like x86's
                        // INT_MIN); the flow element of (int(ty), int(tx)) is element iy * W + ix (grid 1)
                        alive[i] = alive[i] && tx[i] > -1.0f && tx[i] < Wf && ty[i] > -1.0f && ty[i] < Hf;  // hpp:262-266
                        const unsigned bx = __float_as_uint(__fadd_rz(fmaxf(tx[i], 0.0f), 8388608.0f));
                        const unsigned by = __float_as_uint(__fadd_rz(fmaxf(ty[i], 0.0f), 8388608.0f));
                        if (alive[i]) f[i] = __ldg(reinterpret_cast<const float2*>(base_f) + (by * uW + bx - sc_bias));
                        continue;
                    }
                    if (!alive[i]) continue;
                    const int ix = cvt_int(tx[i]), iy = cvt_int(ty[i]);
                    if (ix < 0 || ix >= g.W || iy < 0 || iy >= g.H) {  // hpp:262-266
                        alive[i] = false;
                        continue;
                    }
                    const int fr = cvt_int(div_grid(ty[i], g));
                    const int fc = cvt_int(div_grid(tx[i], g));
                    f[i] = load_flow(base_f, (long long)fr * g.Wf + fc, g);
                }
#pragma unroll
                for (int i = 0; i < kScPix; ++i) {
                    if (!alive[i]) continue;
                    tx[i] = __fadd_rn(tx[i], f[i].x);
                    ty[i] = __fadd_rn(ty[i], f[i].y);
                }
            }
#pragma unroll
            for (int i = 0; i < kScPix; ++i) {
                if (!alive[i]) continue;
                const int ix = cvt_int(tx[i]), iy = cvt_int(ty[i]);
                if (ix < 0 || ix >= g.W || iy < 0 || iy >= g.H) continue;
                const int d = iy * g.W + ix;
                if (general) {
                    atomicMax(win + d, spx[i]);
                } else {
                    dst[d] = uval;
                    const int du = d >> 7;  // occupancy flag of the destination unit (idempotent store)
                    if (of && du != last_flag) {
                        of[du] = 1;
                        last_flag = du;
                    }
                }
            }
        }
        __syncwarp();  // the batch buffer is reused
    }
    span_stamp(span, true);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_warp_gather(const WarpPlan* __restrict__ plan, const uint8_t* __restrict__ new_mask,
                                                         long long new_stride, const uint8_t* __restrict__ state_src,
                                                         uint8_t* __restrict__ state_dst, const int32_t* __restrict__ winner,
                                                         uint8_t* __restrict__ occ_dst, int HW, int n_units, unsigned long long* span) {
    const int t = blockIdx.y;
    const WarpPlan p = plan[t];
    if (p.mode != kWarpScatter || p.uniform_val != 0) return;
    span_stamp(span, false);
    const uint8_t* src = p.src_new ? track_plane(new_mask, new_stride, t) : state_src + (long long)t * HW;
    const int4* win = reinterpret_cast<const int4*>(winner + (long long)t * HW);
    uint32_t* dst = reinterpret_cast<uint32_t*>(state_dst + (long long)t * HW);
    uint8_t* of = occ_dst ? occ_dst + (long long)t * n_units : nullptr;
    const int nq = HW >> 2;
    const int lane = threadIdx.x & 31;
    // one warp iteration = one unit (32 quads)
    for (int u = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); u < n_units; u += gridDim.x * (kThreads / 32)) {
        const int q = u * 32 + lane;
        uint32_t out = 0;
        if (q < nq) {
            const int4 w = win[q];
            const int ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // the zeroed origin of the "no new mask" branch can never be a winner: it is not a source
                uint32_t b = ws[i] >= 0 ? (uint32_t)src[ws[i]] : (uint32_t)p.dflt;
                out |= b << (8 * i);
            }
            dst[q] = out;
        }
        const bool any = __any_sync(0xffffffffu, out != 0u);
        if (of && lane == 0) of[u] = any ? 1 : 0;
    }
    span_stamp(span, true);
}

__global__ void __launch_bounds__(kThreads) k_threshold(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 w = src[i];
        // cv::threshold(.., 1, 255, THRESH_BINARY): v > 1 ? 255 : 0
        w.x = __vcmpgtu4(w.x, 0x01010101u);
        w.y = __vcmpgtu4(w.y, 0x01010101u);
        w.z = __vcmpgtu4(w.z, 0x01010101u);
        w.w = __vcmpgtu4(w.w, 0x01010101u);
        dst[i] = w;
    }
}

}  // namespace

// enough blocks per track to fill the machine at small T, few enough to keep launch tails short at large T
static int plane_blocks(int n16, int T) {
    int bx = (n16 + kThreads - 1) / kThreads;
    int target = max(1, (148 * 8 + T - 1) / T);
    return max(1, min(bx, target));
}

int launch_mask_plan(const MaskSyncArgs& a, cudaStream_t s, bool have_stats) {
    const int T = a.n_tracks;
    const int HW = a.g.HW;
    const int n16 = HW >> 4;
    if (a.new_mask && !have_stats)
        ROFTB_LAUNCH(k_mask_stats, dim3(plane_blocks(n16, T), T), kThreads, 0, s, a.new_mask, a.new_stride, a.ctl, n16, a.stat);
    ROFTB_LAUNCH(k_warp_plan, (T + 127) / 128, 128, 0, s, T, a.ctl, a.stat, a.plan, a.fbuf, a.new_mask, a.new_stride,
                 a.state_src, HW, a.segm_delay, a.fuse);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_mask_init(const MaskSyncArgs& a, cudaStream_t s) {
    const int T = a.n_tracks;
    const int HW = a.g.HW;
    ROFTB_LAUNCH(k_warp_init, dim3(plane_blocks(HW >> 4, T), T), kThreads, 0, s, a.plan, a.new_mask, a.new_stride, a.state_src,
                 a.state_dst, a.winner, a.occ_src, a.occ_dst, HW, a.n_warp_tiles, a.span_clock);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_mask_plan_init(const MaskSyncArgs& a, cudaStream_t s, bool planned) {
    if (!planned && launch_mask_plan(a, s, false)) return -1;
    return launch_mask_init(a, s);
}

int launch_mask_scatter_gather(const MaskSyncArgs& a, cudaStream_t s) {
    const int T = a.n_tracks;
    const int HW = a.g.HW;
    const int nq = HW >> 2;
    int bq = (nq + kThreads - 1) / kThreads;
    int targetq = max(1, (148 * 16 + T - 1) / T);
    bq = max(1, min(bq, targetq));
    int bs = max(1, min((a.n_warp_tiles + 7) / 8, (148 * 16 + T - 1) / T));
    if (!a.g.flow_s16 && a.g.grid == 1 && a.g.scale_mode == 0)
        ROFTB_LAUNCH(k_warp_scatter<true>, dim3(bs, T), kThreads, 0, s, a.g, a.ft, a.plan, a.new_mask, a.new_stride, a.state_src,
                     a.state_dst, a.winner, a.s_list, a.s_n, a.n_list, a.n_n, a.n_warp_tiles, a.occ_dst, a.span_clock ? a.span_clock + 2 : nullptr);
    else
        ROFTB_LAUNCH(k_warp_scatter<false>, dim3(bs, T), kThreads, 0, s, a.g, a.ft, a.plan, a.new_mask, a.new_stride, a.state_src,
                     a.state_dst, a.winner, a.s_list, a.s_n, a.n_list, a.n_n, a.n_warp_tiles, a.occ_dst, a.span_clock ? a.span_clock + 2 : nullptr);
    ROFTB_LAUNCH(k_warp_gather, dim3(bq, T), kThreads, 0, s, a.plan, a.new_mask, a.new_stride, a.state_src, a.state_dst,
                 a.winner, a.occ_dst, HW, a.n_warp_tiles, a.span_clock ? a.span_clock + 4 : nullptr);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_mask_sync(const MaskSyncArgs& a, cudaStream_t s, bool planned) {
    if (launch_mask_plan_init(a, s, planned)) return -1;
    return launch_mask_scatter_gather(a, s);
}

int launch_threshold(const uint8_t* src, uint8_t* dst, size_t n, cudaStream_t s) {
    size_t n16 = n >> 4;
    size_t nb = (n16 + kThreads - 1) / kThreads;
    int bx = nb > (size_t)(148 * 8) ? 148 * 8 : (int)nb;
    ROFTB_LAUNCH(k_threshold, max(bx, 1), kThreads, 0, s, reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), n16);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace roftb
