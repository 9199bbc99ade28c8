// Ordered extraction kernels that share the row-major selection machinery of the velocity path:
//   launch_export_measurement  materialised z / H for bfl-style callers
//        (ImageOpticalFlowMeasurement<T>::measure / getMeasurementMatrix, ...Measurement.hpp:258-283,297-326)
//   launch_masked_points       masked depth de-projection, north-star part 3
//        (CameraMeasurement.cpp:75 -> RobotsIO Camera::point_cloud, restricted to the mask; UPSTREAM-RECALL)
//   launch_masked_depth_l1     inner loop of ROFTFilter::pick_best_alternative (ROFTFilter.cpp:556-566)
//
// Row-major order (the order of cv::findNonZero) is reproduced with two levels of per-warp-tile exclusive
// prefixes: first over the mask candidates (rank -> stride selection), then over the pixels that also pass
// the gates (output position).  Within a warp tile the order is sub-tile, lane, pixel.
#include "roftb_internal.cuh"

namespace roftb {
namespace {

enum { kModeExport = 0, kModePoints = 1, kModeL1 = 2 };

struct ExtractArgs {
    Geom g;
    const uint8_t* mask; long long mask_stride; int thr;
    const float* depth; long long depth_stride;
    const void* flow; long long flow_stride;
    const int32_t* rank_prefix;   // per-warp-tile candidate prefix (needed when stride > 1)
    int32_t* valid_count;         // per-warp-tile valid counts (phase 0: written; phase 1: exclusive prefix)
    int n_warp_tiles;
    int stride;                   // selection stride over the candidates
    double max_depth;
    // export
    double dt, fx, fy, cx, cy;
    int capacity;
    double* z; double* H;
    // points
    double* points;
    // L1
    const float* rendered; long long rendered_stride; int divider; double* err_sum; int32_t* samples;
};

// Evaluates one warp tile: returns the 16-bit mask (bit 4j+i) of pixels that are selected AND pass the gates,
// plus the per-pixel inputs needed by the writer.
template <int MODE>
__device__ __forceinline__ uint32_t eval_tile(const ExtractArgs& a, int t, int wt, int lane, float* dv, float* fxv, float* fyv) {
    const Geom& g = a.g;
    const uint32_t thr4 = (uint32_t)a.thr * 0x01010101u;
    const uint32_t* mq = reinterpret_cast<const uint32_t*>(a.mask + (long long)t * a.mask_stride);
    const float* dp = a.depth + (long long)t * a.depth_stride;
    const int nq = g.HW >> 2;
    uint32_t sel[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int q = wt * 128 + j * 32 + lane;
        const uint32_t m = q < nq ? __ldg(mq + q) : 0u;
        sel[j] = __vcmpgtu4(m, thr4);
    }
    if (a.stride > 1) {
        int r = a.rank_prefix[(long long)t * a.n_warp_tiles + wt];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cnt = __popc(sel[j]) >> 3;
            const int incl = warp_scan_incl(cnt, lane);
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            unsigned rank = (unsigned)(r + incl - cnt);
            uint32_t ns = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if ((sel[j] >> (8 * i)) & 1u) {
                    if (rank % (unsigned)a.stride == 0u) ns |= 0xffu << (8 * i);
                    ++rank;
                }
            }
            sel[j] = ns;
            r += tot;
        }
    }
    uint32_t vmask = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (sel[j] == 0u) continue;
        const int px0 = (wt * 128 + j * 32 + lane) << 2;
        const int v = px0 / g.W;
        const int u0 = px0 - v * g.W;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (!((sel[j] >> (8 * i)) & 1u)) continue;
            const float d = __ldg(dp + px0 + i);
            bool ok;
            if (MODE == kModeExport) {
                const char* fbase = reinterpret_cast<const char*>(a.flow) + (long long)t * a.flow_stride * (g.flow_s16 ? 2 : 4);
                const float2 f = load_flow(fbase, (long long)(v / g.grid) * g.Wf + ((u0 + i) / g.grid), g);
                fxv[4 * j + i] = f.x;
                fyv[4 * j + i] = f.y;
                ok = flow_valid(f.x, f.y) && d > 0.f && (double)d < a.max_depth;
            } else if (MODE == kModePoints) {
                ok = d > 0.f && (double)d < a.max_depth;
            } else {
                const float r = __ldg(a.rendered + (long long)t * a.rendered_stride +
                                      (long long)(v / a.divider) * (g.W / a.divider) + (u0 + i) / a.divider);
                fxv[4 * j + i] = r;
                ok = d > 0.f && (double)d < 2.0 && r != 0.0f;
            }
            dv[4 * j + i] = d;
            if (ok) vmask |= 1u << (4 * j + i);
        }
    }
    return vmask;
}

template <int MODE, int PHASE>
__global__ void __launch_bounds__(kThreads) k_extract(ExtractArgs a) {
    const int t = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Geom& g = a.g;
    double l1_err = 0.0;
    int l1_n = 0;
    for (int wt = blockIdx.x * (kThreads / 32) + warp; wt < a.n_warp_tiles; wt += gridDim.x * (kThreads / 32)) {
        float dv[16], fxv[16], fyv[16];
        const uint32_t vmask = eval_tile<MODE>(a, t, wt, lane, dv, fxv, fyv);
        if (MODE == kModeL1) {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if ((vmask >> k) & 1u) {
                    l1_err += (double)fabsf(dv[k] - fxv[k]);  // std::abs(float - float) accumulated in double
                    ++l1_n;
                }
            continue;
        }
        if (PHASE == 0) {
            const int c = warp_sum(__popc(vmask));
            if (lane == 0) a.valid_count[(long long)t * a.n_warp_tiles + wt] = c;
            continue;
        }
        // PHASE 1: ordered write
        int base = a.valid_count[(long long)t * a.n_warp_tiles + wt];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cnt = __popc((vmask >> (4 * j)) & 0xfu);
            const int incl = warp_scan_incl(cnt, lane);
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            int pos = base + incl - cnt;
            const int px0 = (wt * 128 + j * 32 + lane) << 2;
            const int v = px0 / g.W;
            const int u0 = px0 - v * g.W;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!((vmask >> (4 * j + i)) & 1u)) continue;
                if (pos < a.capacity) {
                    const double d = (double)dv[4 * j + i];
                    const double uu = (double)(u0 + i) - a.cx, vv = (double)v - a.cy;
                    if (MODE == kModeExport) {
                        // hpp:272-282, FP64 like the reference
                        double* zr = a.z + ((long long)t * a.capacity + pos) * 2;
                        double* hr = a.H + ((long long)t * a.capacity + pos) * 12;
                        zr[0] = (double)fxv[4 * j + i];
                        zr[1] = (double)fyv[4 * j + i];
                        hr[0] = a.fx / d * a.dt;
                        hr[1] = 0.0;
                        hr[2] = -uu / d * a.dt;
                        hr[3] = -uu * vv / a.fy * a.dt;
                        hr[4] = (a.fx + uu * uu / a.fx) * a.dt;
                        hr[5] = -vv * a.fx / a.fy * a.dt;
                        hr[6] = 0.0;
                        hr[7] = a.fy / d * a.dt;
                        hr[8] = -vv / d * a.dt;
                        hr[9] = -(a.fy + vv * vv / a.fy) * a.dt;
                        hr[10] = vv * uu / a.fx * a.dt;
                        hr[11] = uu * a.fy / a.fx * a.dt;
                    } else {
                        double* pr = a.points + ((long long)t * a.capacity + pos) * 3;
                        pr[0] = uu * d / a.fx;
                        pr[1] = vv * d / a.fy;
                        pr[2] = d;
                    }
                }
                ++pos;
            }
            base += tot;
        }
    }
    if (MODE == kModeL1) {
        l1_err = warp_sum(l1_err);
        l1_n = warp_sum(l1_n);
        if (lane == 0 && l1_n) {
            atomicAdd(a.err_sum + t, l1_err);
            atomicAdd(a.samples + t, l1_n);
        }
    }
}

// ---- buffered features of the render-and-compare pose test, compacted -------------------------------------------
// ROFTFilter::pick_best_alternative only ever looks at every second non-zero pixel of the (buffered) segmentation, in
// findNonZero order, and at the depth under it (ROFTFilter.cpp:556-566).  Instead of keeping copies of both planes
// (5 bytes per pixel of the frame), the features of a track are that list itself: entry k = (linear pixel index, depth
// bits) of coordinate 2k, depth bits 0 when the depth gate 0 < d < 2 fails (a valid depth is never 0).  ~0.9 MB per
// track at 25 % coverage instead of 4.6 MB, and the L1 of both rendered alternatives is one pass over it.
__global__ void __launch_bounds__(kThreads) k_or_features(Geom g, const UkfOp* __restrict__ ops, int max_ops, int bits,
                                                          const uint8_t* __restrict__ mask, long long mask_stride, int thr,
                                                          const float* __restrict__ depth, long long depth_stride,
                                                          const int32_t* __restrict__ rank_prefix, const int32_t* __restrict__ total,
                                                          int n_warp_tiles, uint2* __restrict__ feat, long long feat_stride,
                                                          int32_t* __restrict__ n_feat) {
    const int t = blockIdx.y;
    if ((ops[(long long)t * max_ops].pad & bits) == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) n_feat[t] = (total[t] + 1) >> 1;
    const uint32_t thr4 = (uint32_t)thr * 0x01010101u;
    const uint32_t* mq = reinterpret_cast<const uint32_t*>(mask + (long long)t * mask_stride);
    const float* dp = depth + (long long)t * depth_stride;
    uint2* out = feat + (long long)t * feat_stride;
    const int nq = g.HW >> 2;
    for (int wt = blockIdx.x * (kThreads / 32) + warp; wt < n_warp_tiles; wt += gridDim.x * (kThreads / 32)) {
        int r = rank_prefix[(long long)t * n_warp_tiles + wt];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = wt * 128 + j * 32 + lane;
            const uint32_t sel = q < nq ? __vcmpgtu4(__ldg(mq + q), thr4) : 0u;
            const int cnt = __popc(sel) >> 3;
            const int incl = warp_scan_incl(cnt, lane);
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            int rank = r + incl - cnt;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!((sel >> (8 * i)) & 1u)) continue;
                if ((rank & 1) == 0) {
                    const int px = (q << 2) + i;
                    const float d = __ldg(dp + px);
                    const bool ok = d > 0.f && (double)d < 2.0;
                    out[rank >> 1] = make_uint2((uint32_t)px, ok ? __float_as_uint(d) : 0u);
                }
                ++rank;
            }
            r += tot;
        }
    }
}

// stage -> snapshot of the flagged tracks
__global__ void __launch_bounds__(kThreads) k_or_feat_copy(const UkfOp* __restrict__ ops, int max_ops, int bits,
                                                           const uint2* __restrict__ src, const int32_t* __restrict__ n_src,
                                                           uint2* __restrict__ dst, int32_t* __restrict__ n_dst, long long stride) {
    const int t = blockIdx.y;
    if ((ops[(long long)t * max_ops].pad & bits) == 0) return;
    const int n = n_src[t];
    if (blockIdx.x == 0 && threadIdx.x == 0) n_dst[t] = n;
    const uint4* s4 = reinterpret_cast<const uint4*>(src + (long long)t * stride);
    uint4* d4 = reinterpret_cast<uint4*>(dst + (long long)t * stride);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n + 1) / 2; i += gridDim.x * blockDim.x) d4[i] = s4[i];
}

// masked depth L1 of BOTH rendered alternatives over the feature list of every track with a pending test: one block per
// track, fixed reduction order (deterministic).  err / samples laid out [alternative][track].
__global__ void __launch_bounds__(1024) k_or_l1(int n_tracks, const int32_t* __restrict__ resume, const uint2* __restrict__ feat,
                                                const int32_t* __restrict__ n_feat, long long feat_stride,
                                                const float* __restrict__ rendered, long long tile, int divider, int W,
                                                double* __restrict__ err, int32_t* __restrict__ samples) {
    const int t = blockIdx.x;
    __shared__ double s_e[2][32];
    __shared__ int s_n[2][32];
    double e0 = 0.0, e1 = 0.0;
    int n0 = 0, n1 = 0;
    if (resume[t] > 0) {
        const uint2* f = feat + (long long)t * feat_stride;
        const float* ra = rendered + (long long)t * tile;
        const float* rb = rendered + ((long long)n_tracks + t) * tile;
        const int n = n_feat[t], wt = W / divider;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint2 en = f[i];
            if (en.y == 0u) continue;
            const float d = __uint_as_float(en.y);
            const int v = (int)(en.x / (unsigned)W), u = (int)(en.x - (unsigned)v * (unsigned)W);
            const long long ri = (long long)(v / divider) * wt + u / divider;
            const float a = __ldg(ra + ri), b = __ldg(rb + ri);
            if (a != 0.0f) { e0 += (double)fabsf(d - a); ++n0; }   // std::abs(float - float) accumulated in double
            if (b != 0.0f) { e1 += (double)fabsf(d - b); ++n1; }
        }
    }
    e0 = warp_sum(e0); e1 = warp_sum(e1);
    n0 = warp_sum(n0); n1 = warp_sum(n1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_e[0][warp] = e0; s_e[1][warp] = e1; s_n[0][warp] = n0; s_n[1][warp] = n1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double e = 0.0;
        int n = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { e += s_e[threadIdx.x][w]; n += s_n[threadIdx.x][w]; }
        err[(long long)threadIdx.x * n_tracks + t] = e;
        samples[(long long)threadIdx.x * n_tracks + t] = n;
    }
}

ExtractArgs base_args(const SelectArgs& s) {
    ExtractArgs a;
    memset(&a, 0, sizeof(a));
    a.g = s.g;
    a.mask = s.mask; a.mask_stride = s.mask_stride; a.thr = s.thr;
    a.depth = s.depth; a.depth_stride = s.depth_stride;
    a.flow = s.flow; a.flow_stride = s.flow_stride;
    a.rank_prefix = s.wt_count;
    a.valid_count = s.wt_count2;
    a.n_warp_tiles = (s.g.HW + kWarpTilePx - 1) / kWarpTilePx;
    a.stride = 1;
    a.capacity = 0;
    return a;
}

}  // namespace

int launch_export_measurement(const SelectArgs& s, double dt, double fx, double fy, double cx, double cy, int capacity,
                              double* z, double* H, int32_t* n_valid, cudaStream_t st) {
    ExtractArgs a = base_args(s);
    a.stride = s.g.stride;
    a.max_depth = s.g.max_depth;
    a.dt = dt; a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy;
    a.capacity = capacity; a.z = z; a.H = H;
    if (a.stride > 1 && launch_mask_rank(s.mask, s.mask_stride, s.thr, s.g.HW, s.n_items, s.wt_count, nullptr, nullptr, st)) return -1;
    const int bx = max(1, min(a.n_warp_tiles / (kThreads / 32), 148 * 4));
    ROFTB_LAUNCH((k_extract<kModeExport, 0>), dim3(bx, s.n_items), kThreads, 0, st, a);
    if (launch_wt_scan(s.wt_count2, a.n_warp_tiles, s.n_items, n_valid, st)) return -1;
    ROFTB_LAUNCH((k_extract<kModeExport, 1>), dim3(bx, s.n_items), kThreads, 0, st, a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_masked_points(const SelectArgs& s, double max_depth, double fx, double fy, double cx, double cy, int capacity,
                         double* points, int32_t* count, cudaStream_t st) {
    ExtractArgs a = base_args(s);
    a.max_depth = max_depth;
    a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy;
    a.capacity = capacity; a.points = points;
    const int bx = max(1, min(a.n_warp_tiles / (kThreads / 32), 148 * 4));
    ROFTB_LAUNCH((k_extract<kModePoints, 0>), dim3(bx, s.n_items), kThreads, 0, st, a);
    if (launch_wt_scan(s.wt_count2, a.n_warp_tiles, s.n_items, count, st)) return -1;
    ROFTB_LAUNCH((k_extract<kModePoints, 1>), dim3(bx, s.n_items), kThreads, 0, st, a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_masked_depth_l1(const SelectArgs& s, const float* rendered, long long rendered_stride, int divider, double* err_sum,
                           int32_t* samples, cudaStream_t st) {
    ExtractArgs a = base_args(s);
    a.stride = 2;  // every 2nd non-zero pixel (ROFTFilter.cpp:556)
    a.rendered = rendered; a.rendered_stride = rendered_stride; a.divider = divider;
    a.err_sum = err_sum; a.samples = samples;
    if (launch_mask_rank(s.mask, s.mask_stride, s.thr, s.g.HW, s.n_items, s.wt_count, nullptr, nullptr, st)) return -1;
    const int bx = max(1, min(a.n_warp_tiles / (kThreads / 32), 148 * 4));
    ROFTB_LAUNCH((k_extract<kModeL1, 0>), dim3(bx, s.n_items), kThreads, 0, st, a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// features of the flagged tracks (UkfOp::pad & bits) from a raw mask state plane (thr = 1) and a depth plane; wt_count /
// total are scratch ([n][n_warp_tiles], [n])
int launch_or_features(const Geom& g, int n_tracks, const UkfOp* ops, int max_ops, int bits, const uint8_t* mask, long long mask_stride,
                       int thr, const float* depth, long long depth_stride, int32_t* wt_count, int32_t* total, uint2* feat,
                       long long feat_stride, int32_t* n_feat, cudaStream_t st) {
    const int n_warp_tiles = (g.HW + kWarpTilePx - 1) / kWarpTilePx;
    if (launch_mask_rank(mask, mask_stride, thr, g.HW, n_tracks, wt_count, total, nullptr, st)) return -1;
    const int bx = max(1, min(n_warp_tiles / (kThreads / 32), max(1, 148 * 8 / n_tracks)));
    ROFTB_LAUNCH(k_or_features, dim3(bx, n_tracks), kThreads, 0, st, g, ops, max_ops, bits, mask, mask_stride, thr, depth, depth_stride,
                 wt_count, total, n_warp_tiles, feat, feat_stride, n_feat);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_or_feat_copy(int n_tracks, const UkfOp* ops, int max_ops, int bits, const uint2* src, const int32_t* n_src, uint2* dst,
                        int32_t* n_dst, long long stride, cudaStream_t st) {
    ROFTB_LAUNCH(k_or_feat_copy, dim3(max(1, 148 * 4 / n_tracks), n_tracks), kThreads, 0, st, ops, max_ops, bits, src, n_src, dst, n_dst,
                 stride);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_or_l1(int n_tracks, const int32_t* resume, const uint2* feat, const int32_t* n_feat, long long feat_stride,
                 const float* rendered, long long tile, int divider, int W, double* err, int32_t* samples, cudaStream_t st) {
    ROFTB_LAUNCH(k_or_l1, n_tracks, 1024, 0, st, n_tracks, resume, feat, n_feat, feat_stride, rendered, tile, divider, W, err, samples);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace roftb
