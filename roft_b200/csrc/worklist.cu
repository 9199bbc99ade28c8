// Worklists of the non-empty 128-pixel units of a byte plane, and row-major rank bases.
//
// Masks cover a compact fraction of the frame, so every streaming kernel iterates over the list of non-empty UNITS of
// its track (a unit = 128 consecutive pixels = one quad per lane of a warp) instead of the whole plane.  Two sources:
//   * a plane that arrives from outside (a newly delivered mask, an operator argument) is read once by
//     k_tile_count / k_unit_flags;
//   * the synchronised mask state never is: every kernel that WRITES it also sets one occupancy flag per unit it
//     touches (mask_sync.cu, velocity_track.cu), and the consumers compact those flags.
// The rank kernels serve cv::findNonZero's row-major order (ImageOpticalFlowMeasurement.hpp:233-237) for the ordered
// export / point-cloud operators (extract.cu).
#include "roftb_internal.cuh"

namespace roftb {
namespace {

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
                                                        const int32_t* __restrict__ active, int active_stride,
                                                        MaskStat* __restrict__ stat) {
    const int t = blockIdx.y;
    if (active && !active[(long long)t * active_stride]) return;
    // optional: non-zero count / min / max of the plane in the same read (emptiness and single-valuedness of a newly
    // delivered mask, ImageSegmentationOFAidedSource.hpp:186-197)
    int st_nnz = 0;
    uint32_t st_min = 0xffffffffu, st_max = 0u;
    auto stats16 = [&](const uint4& w) {
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t nz = __vcmpne4(ws[k], 0u);
            st_nnz += __popc(nz) >> 3;
            st_min = __vminu4(st_min, ws[k] | ~nz);
            st_max = __vmaxu4(st_max, ws[k]);
        }
    };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t thr4 = (uint32_t)thr * 0x01010101u;
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
        if (stat) {
            stats16(wa);
            stats16(wb);
        }
        if ((lane & 7) == 0) {
            const int ua = b * 4 + (lane >> 3), ub = b2 * 4 + (lane >> 3);
            if (ua < n_units) wt_count[(long long)t * n_units + ua] = (int)ca;
            if (b2 < n_blk && ub < n_units) wt_count[(long long)t * n_units + ub] = (int)cb;
        }
    }
    if (stat) {
        int mn = min(min(st_min & 0xff, (st_min >> 8) & 0xff), min((st_min >> 16) & 0xff, st_min >> 24));
        int mx = max(max(st_max & 0xff, (st_max >> 8) & 0xff), max((st_max >> 16) & 0xff, st_max >> 24));
        st_nnz = warp_sum(st_nnz);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0 && st_nnz > 0) {
            atomicAdd(&stat[t].nnz, st_nnz);
            atomicMin(&stat[t].vmin, mn);
            atomicMax(&stat[t].vmax, mx);
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

// ---- occupancy flags ------------------------------------------------------------------------------
// one byte per unit: 1 iff the unit holds a non-zero byte.  Used for planes that come from outside (operator mode);
// the mask state carries its flags along (see the header comment).
__global__ void __launch_bounds__(kThreads) k_unit_flags(const uint8_t* __restrict__ plane, long long stride, int HW, int n_units,
                                                        uint8_t* __restrict__ flags) {
    const int t = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4* m4 = reinterpret_cast<const uint4*>(plane + (long long)t * stride);
    const int n16 = HW >> 4;
    const int n_blk = (n_units + 3) >> 2;  // 512-px blocks: one 128-bit load per lane, 8 lanes per unit
    const unsigned gmask = 0xffu << (8 * (lane >> 3));
    for (int b = blockIdx.x * (kThreads / 32) + warp; b < n_blk; b += gridDim.x * (kThreads / 32)) {
        const int q = b * 32 + lane;
        const uint4 w = q < n16 ? ld_nc_u4(m4 + q) : make_uint4(0u, 0u, 0u, 0u);
        const unsigned any = __reduce_or_sync(gmask, w.x | w.y | w.z | w.w);
        const int u = b * 4 + (lane >> 3);
        if ((lane & 7) == 0 && u < n_units) flags[(long long)t * n_units + u] = any ? 1 : 0;
    }
}

// one block per track: list <- ids of the flagged units (ascending), n[t] <- their number.  `plan` (optional) restricts
// the work to the tracks whose mask is scattered from the STATE plane by the stand-alone scatter kernel.
__global__ void __launch_bounds__(kCompactThreads) k_flag_list(const uint8_t* __restrict__ flags, int n_units,
                                                               int32_t* __restrict__ wt_list, int32_t* __restrict__ wt_n,
                                                               const WarpPlan* __restrict__ plan) {
    const int t = blockIdx.x;
    if (plan) {
        const WarpPlan& p = plan[t];
        if (p.mode != kWarpScatter || p.fused || p.src_new) return;
    }
    __shared__ int sh[kCompactThreads / 32];
    __shared__ int carry;
    const uint8_t* f = flags + (long long)t * n_units;
    int32_t* list = wt_list + (long long)t * n_units;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_units; base += kCompactThreads) {
        const int i = base + threadIdx.x;
        const int v = (i < n_units && f[i]) ? 1 : 0;
        const int incl = warp_scan_incl(v, lane);
        if (lane == 31) sh[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += sh[w];
        const int c = carry;
        if (v) list[c + woff + incl - 1] = i;
        __syncthreads();
        if (threadIdx.x == kCompactThreads - 1) carry = c + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) wt_n[t] = carry;
}

// Longest-first launch order of the velocity kernel's clusters from the occupancy flags of a mask plane: flagged units
// per track, then the counting sort the velocity kernel itself runs at the end of a launch (velocity_track.cu,
// write_next_order - same buckets).  Used after a step that delivered new masks: the order the velocity kernel derives
// from ITS worklists describes the masks before the delivery, and one frame-filling mask scheduled last doubles the
// span of the next launch (measured on B200: 1.42 ms instead of 0.75 ms at 256 tracks).
__global__ void __launch_bounds__(kThreads) k_order_from_flags(const uint8_t* __restrict__ flags, int n_units, int n_tracks,
                                                               int32_t* __restrict__ units, uint32_t* __restrict__ ticket,
                                                               int32_t* __restrict__ order) {
    __shared__ int s_bucket[129];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * (kThreads / 32) + warp;  // one warp per track
    if (t < n_tracks) {
        const uint8_t* f = flags + (long long)t * n_units;
        int n = 0;
        if ((n_units & 15) == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0) {  // 16 flags per load
            const uint4* f4 = reinterpret_cast<const uint4*>(f);
            auto nzb = [](uint32_t x) { return __popc((((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u); };
#pragma unroll 4
            for (int i = lane; i < (n_units >> 4); i += 32) {
                const uint4 v = __ldcg(f4 + i);
                n += nzb(v.x) + nzb(v.y) + nzb(v.z) + nzb(v.w);
            }
        } else {
            for (int i = lane; i < n_units; i += 32) n += f[i] ? 1 : 0;
        }
        n = warp_sum(n);
        if (lane == 0) units[t] = n;
    }
    // the last block to finish sorts (ticket returns to zero for the next launch)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned v = atomicAdd(ticket, 1u);
        s_last = v == gridDim.x - 1;
        if (s_last) *ticket = 0u;
    }
    __syncthreads();
    if (!s_last || warp != 0) return;
    __threadfence();
    constexpr int NB = 128;
    for (int i = lane; i <= NB; i += 32) s_bucket[i] = 0;
    __syncwarp();
    const int shift = 32 - __clz(max(n_units, 1) / NB + 1);
    for (int k = lane; k < n_tracks; k += 32) atomicAdd(&s_bucket[NB - min(NB - 1, __ldcg(units + k) >> shift)], 1);
    __syncwarp();
    if (lane == 0)
        for (int i = 1; i <= NB; ++i) s_bucket[i] += s_bucket[i - 1];
    __syncwarp();
    // (stable within a bucket is not required: equal-sized tracks are interchangeable for the schedule)
    for (int k = lane; k < n_tracks; k += 32) order[atomicAdd(&s_bucket[NB - 1 - min(NB - 1, __ldcg(units + k) >> shift)], 1)] = k;
}

}  // namespace

int launch_order_from_flags(const uint8_t* flags, int n_units, int n_tracks, int32_t* units_tmp, uint32_t* ticket, int32_t* order,
                            cudaStream_t s) {
    ROFTB_LAUNCH(k_order_from_flags, (n_tracks + kThreads / 32 - 1) / (kThreads / 32), kThreads, 0, s, flags, n_units, n_tracks, units_tmp,
                 ticket, order);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_unit_flags(const uint8_t* plane, long long stride, int HW, int n_items, uint8_t* flags, cudaStream_t s) {
    const int n_units = (HW + kUnitPx - 1) / kUnitPx;
    const int n_block_tiles = (HW + kBlockTilePx - 1) / kBlockTilePx;
    const int bx = max(1, min(n_block_tiles, (148 * 8 + n_items - 1) / n_items));
    ROFTB_LAUNCH(k_unit_flags, dim3(bx, n_items), kThreads, 0, s, plane, stride, HW, n_units, flags);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_flag_list(const uint8_t* flags, int n_units, int n_items, int32_t* wt_list, int32_t* wt_n, const WarpPlan* plan,
                     cudaStream_t s) {
    ROFTB_LAUNCH(k_flag_list, n_items, kCompactThreads, 0, s, flags, n_units, wt_list, wt_n, plan);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

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
                     int32_t* wt_n, const int32_t* active, int active_stride, cudaStream_t s, MaskStat* stat) {
    const int n_units = (HW + kUnitPx - 1) / kUnitPx;
    const int n_block_tiles = (HW + kBlockTilePx - 1) / kBlockTilePx;
    const int bx = max(1, min(n_block_tiles, (148 * 8 + n_items - 1) / n_items));
    ROFTB_LAUNCH(k_tile_count, dim3(bx, n_items), kThreads, 0, s, plane, stride, thr, HW, n_units, wt_count, active,
                 active_stride, stat);
    ROFTB_LAUNCH(k_tile_compact, n_items, kCompactThreads, 0, s, wt_count, wt_list, wt_n, n_units, active, active_stride);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace roftb
