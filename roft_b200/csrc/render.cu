// render.cu - depth rasteriser for the pose outlier rejection.
//
// Replaces SICAD::superimpose(poses, cam_x = 0, cam_o = identity, ..., depth) (SICAD.cpp:924-1066) with the fragment
// shader's linearised depth output (shader_model.frag:33-52) for the use ROFTFilter::pick_best_alternative makes of it
// (ROFTFilter.cpp:496-541): n_items object poses -> n_items depth tiles of (W / divider) x (H / divider) pixels, 0 where
// no surface is hit.  An OpenGL context is neither available nor wanted here: the pipeline is restated as three kernels
//   vertex  : model transform (the float model matrix of glm::rotate + translation, built on the host), pinhole
//             projection equal to SICAD's projection matrix followed by the viewport transform and the cv::flip
//             (SICAD.cpp:1634-1637: u = fx X / Z + cx, v = fy Y / Z + cy, pixel centres at +0.5), window depth
//             z_win = (z_ndc + 1) / 2 with near 0.001, far 1000 in FP32; positions snapped to 1/256 pixel like the
//             hardware rasteriser does
//   raster  : one warp per triangle, lanes over the pixels of its bounding box, exact 64-bit edge functions on the snapped
//             positions (inclusive edges: no cracks between triangles), z_win interpolated with the integer barycentrics
//             in FP64 and rounded to FP32, depth test = atomicMin on the bits of the (positive) float
//   resolve : linearize_depth of the surviving z_win (shader_model.frag:39-46), 0 for untouched pixels
// Triangles with a vertex at or behind the near plane are dropped (no clipping: tracked objects are 0.3 - 2 m away).
// GL's own rasterisation is implementation-defined in the last bits, so this row's parity is against oracle/ (same
// restatement in numpy) with a tolerance, not bit-exact: DESIGN.md 4.4.
#include "roftb_internal.cuh"

namespace roftb {

namespace {

constexpr float kNear = 0.001f, kFar = 1000.0f;
constexpr int kSub = 256;  // sub-pixel positions per pixel

struct RVertex {
    int x, y;   // window position in 1/256 px (x right, y down, origin = top-left corner of the tile)
    float z;    // window depth in [0, 1]
    int ok;     // in front of the near plane and finite
};

__global__ void __launch_bounds__(kThreads) k_render_vertices(RenderArgs a, RVertex* __restrict__ out) {
    const int item = blockIdx.y;
    const float* m = a.model + (long long)item * 12;
    // optional per-track scale of the shared mesh (batched tracks of differently sized objects): items are [alternative][track]
    const float* sc = a.scale ? a.scale + 3 * (item % a.n_scale) : nullptr;
    const float sx = sc ? sc[0] : 1.f, sy = sc ? sc[1] : 1.f, sz = sc ? sc[2] : 1.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_vertices; i += gridDim.x * blockDim.x) {
        const float x = __fmul_rn(a.vertices[3 * i], sx), y = __fmul_rn(a.vertices[3 * i + 1], sy), z = __fmul_rn(a.vertices[3 * i + 2], sz);
        // p = R v + t, IEEE FP32 without contraction
        const float X = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), m[9]);
        const float Y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[3], x), __fmul_rn(m[4], y)), __fmul_rn(m[5], z)), m[10]);
        const float Z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[6], x), __fmul_rn(m[7], y)), __fmul_rn(m[8], z)), m[11]);
        RVertex v;
        v.ok = (Z > kNear && Z < kFar) ? 1 : 0;
        const float iz = __fdiv_rn(1.0f, v.ok ? Z : 1.0f);
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(a.fx, X), iz), a.cx);
        const float w = __fadd_rn(__fmul_rn(__fmul_rn(a.fy, Y), iz), a.cy);
        // z_ndc = (f + n) / (f - n) - 2 f n / ((f - n) Z)
        const float A = __fdiv_rn(kFar + kNear, kFar - kNear), B = __fdiv_rn(2.0f * kFar * kNear, kFar - kNear);
        const float zn = __fadd_rn(A, -__fmul_rn(B, iz));
        v.z = __fadd_rn(__fmul_rn(0.5f, zn), 0.5f);
        const float lim = 1.0e6f;  // keeps the 64-bit edge functions far from overflow
        const float us = fminf(fmaxf(__fmul_rn(u, (float)kSub), -lim * kSub), lim * kSub);
        const float ws = fminf(fmaxf(__fmul_rn(w, (float)kSub), -lim * kSub), lim * kSub);
        if (!(us == us) || !(ws == ws)) v.ok = 0;
        v.x = __float2int_rn(us);
        v.y = __float2int_rn(ws);
        out[(long long)item * a.n_vertices + i] = v;
    }
}

__global__ void __launch_bounds__(kThreads) k_render_triangles(RenderArgs a, const RVertex* __restrict__ vtx, uint32_t* __restrict__ zbuf) {
    const int item = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const RVertex* V = vtx + (long long)item * a.n_vertices;
    uint32_t* zb = zbuf + (long long)item * a.w * a.h;
    for (int f = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); f < a.n_faces; f += gridDim.x * (kThreads / 32)) {
        const int i0 = a.faces[3 * f], i1 = a.faces[3 * f + 1], i2 = a.faces[3 * f + 2];
        if ((unsigned)i0 >= (unsigned)a.n_vertices || (unsigned)i1 >= (unsigned)a.n_vertices || (unsigned)i2 >= (unsigned)a.n_vertices)
            continue;
        const RVertex v0 = V[i0];
        RVertex v1 = V[i1], v2 = V[i2];
        if (!(v0.ok && v1.ok && v2.ok)) continue;
        long long area = (long long)(v1.x - v0.x) * (v2.y - v0.y) - (long long)(v1.y - v0.y) * (v2.x - v0.x);
        if (area == 0) continue;
        if (area < 0) {  // same orientation for every triangle: no face culling (SICAD never enables it)
            const RVertex t = v1;
            v1 = v2;
            v2 = t;
            area = -area;
        }
        // pixels whose centre (k + 0.5) can lie inside: k*256 + 128 in [min, max]
        const int xmin = max(0, (min(v0.x, min(v1.x, v2.x)) - kSub / 2 + kSub - 1) >> 8);
        const int xmax = min(a.w - 1, (max(v0.x, max(v1.x, v2.x)) - kSub / 2) >> 8);
        const int ymin = max(0, (min(v0.y, min(v1.y, v2.y)) - kSub / 2 + kSub - 1) >> 8);
        const int ymax = min(a.h - 1, (max(v0.y, max(v1.y, v2.y)) - kSub / 2) >> 8);
        if (xmin > xmax || ymin > ymax) continue;
        const int bw = xmax - xmin + 1;
        const long long npx = (long long)bw * (ymax - ymin + 1);
        const double inv_area = 1.0 / (double)area;
        for (long long p = lane; p < npx; p += 32) {
            const int py = ymin + (int)(p / bw), px = xmin + (int)(p % bw);
            const long long cx = (long long)px * kSub + kSub / 2, cy = (long long)py * kSub + kSub / 2;
            const long long w0 = (v2.x - v1.x) * (cy - v1.y) - (v2.y - v1.y) * (cx - v1.x);
            const long long w1 = (v0.x - v2.x) * (cy - v2.y) - (v0.y - v2.y) * (cx - v2.x);
            const long long w2 = area - w0 - w1;
            if (w0 < 0 || w1 < 0 || w2 < 0) continue;
            const double zd = ((double)w0 * (double)v0.z + (double)w1 * (double)v1.z + (double)w2 * (double)v2.z) * inv_area;
            const float zf = (float)zd;
            if (zf >= 0.0f && zf <= 1.0f) atomicMin(zb + (long long)py * a.w + px, __float_as_uint(zf));
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_render_resolve(const uint32_t* __restrict__ zbuf, float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t b = zbuf[i];
        float d = 0.0f;
        if (b != 0xffffffffu) {
            // linearize_depth (shader_model.frag:39-46)
            const float z = __fadd_rn(__fmul_rn(__uint_as_float(b), 2.0f), -1.0f);
            d = __fdiv_rn(2.0f * kNear * kFar, __fadd_rn(kFar + kNear, -__fmul_rn(z, kFar - kNear)));
        }
        out[i] = d;
    }
}

// pick_best_alternative's decision (ROFTFilter.cpp:568-583): likelihood = mean error / gain, DBL_MAX without samples;
// the second alternative wins iff likelihood[0] > 2 likelihood[1]
// err / samples are laid out [alternative][track] (one L1 launch per alternative), likelihoods [track][2]
__global__ void k_pick_best(int n, const double* __restrict__ err, const int32_t* __restrict__ samples, double gain,
                            int32_t* __restrict__ selected, double* __restrict__ likelihoods) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double l[2];
    for (int k = 0; k < 2; ++k) {
        const int s = samples[k * n + t];
        l[k] = s == 0 ? 1.7976931348623157e308 : (err[k * n + t] / (double)s) / gain;
        if (likelihoods) likelihoods[2 * t + k] = l[k];
    }
    selected[t] = l[0] > 2.0 * l[1] ? 1 : 0;
}


// In-loop variant of the host-side pose conversion (roftb_api.cu: state_to_pose7 + model_matrix): the two candidate means
// of every track with a pending render-and-compare test -> model matrices laid out [alternative][track]; tracks without
// one are placed behind the camera (nothing is rendered, no samples, first alternative kept).
__global__ void k_or_models(int n_tracks, const int32_t* __restrict__ resume, const double* __restrict__ cand_mean,
                            float* __restrict__ model) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n_tracks) return;
    const int k = i / n_tracks, t = i - k * n_tracks;
    float* out = model + (long long)i * 12;
    if (resume[t] <= 0) {
        for (int j = 0; j < 12; ++j) out[j] = 0.f;
        out[0] = out[4] = out[8] = 1.f;
        out[11] = -1.f;
        return;
    }
    const double* m = cand_mean + ((long long)t * 2 + k) * 13;
    // Eigen::AngleAxisd(Quaterniond) (ROFTFilter.cpp:518-524)
    const double w = m[9], x = m[10], y = m[11], z = m[12];
    double n = sqrt(x * x + y * y + z * z), angle = 0.0, axd[3] = {1.0, 0.0, 0.0};
    if (n != 0.0) {
        angle = 2.0 * atan2(n, fabs(w));
        if (w < 0.0) n = -n;
        axd[0] = x / n; axd[1] = y / n; axd[2] = z / n;
    }
    // glm::rotate(I, float(angle), float(axis)) (SICAD.cpp:604-607)
    const float ang = (float)angle;
    float ax = (float)axd[0], ay = (float)axd[1], az = (float)axd[2];
    const float c = cosf(ang), s = sinf(ang);
    const float nn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az)));
    if (nn > 0.f) { ax = __fdiv_rn(ax, nn); ay = __fdiv_rn(ay, nn); az = __fdiv_rn(az, nn); }
    const float omc = __fadd_rn(1.f, -c);
    const float tx = __fmul_rn(omc, ax), ty = __fmul_rn(omc, ay), tz = __fmul_rn(omc, az);
    out[0] = __fadd_rn(c, __fmul_rn(tx, ax));
    out[1] = __fadd_rn(__fmul_rn(ty, ax), -__fmul_rn(s, az));
    out[2] = __fadd_rn(__fmul_rn(tz, ax), __fmul_rn(s, ay));
    out[3] = __fadd_rn(__fmul_rn(tx, ay), __fmul_rn(s, az));
    out[4] = __fadd_rn(c, __fmul_rn(ty, ay));
    out[5] = __fadd_rn(__fmul_rn(tz, ay), -__fmul_rn(s, ax));
    out[6] = __fadd_rn(__fmul_rn(tx, az), -__fmul_rn(s, ay));
    out[7] = __fadd_rn(__fmul_rn(ty, az), __fmul_rn(s, ax));
    out[8] = __fadd_rn(c, __fmul_rn(tz, az));
    out[9] = (float)m[6]; out[10] = (float)m[7]; out[11] = (float)m[8];
}

// correction = best_alternative (ROFTFilter.cpp:670-673): candidate 1 replaces the belief where it was selected
__global__ void k_or_select(int n_tracks, const int32_t* __restrict__ resume, const int32_t* __restrict__ selected,
                            const double* __restrict__ cand_mean, const double* __restrict__ cand_cov, double* __restrict__ mean,
                            double* __restrict__ cov) {
    const int t = blockIdx.x;
    if (resume[t] <= 0 || selected[t] != 1) return;
    for (int i = threadIdx.x; i < 13; i += blockDim.x) mean[(long long)t * 13 + i] = cand_mean[(long long)t * 26 + 13 + i];
    for (int i = threadIdx.x; i < 144; i += blockDim.x) cov[(long long)t * 144 + i] = cand_cov[(long long)t * 288 + 144 + i];
}

}  // namespace

int launch_render_depth(const RenderArgs& a, void* vertex_scratch, uint32_t* zbuf, float* out, cudaStream_t s) {
    if (a.n_items <= 0 || a.n_vertices <= 0 || a.n_faces <= 0) return -1;
    RVertex* vt = reinterpret_cast<RVertex*>(vertex_scratch);
    const long long npx = (long long)a.n_items * a.w * a.h;
    if (cudaMemsetAsync(zbuf, 0xFF, (size_t)npx * 4, s) != cudaSuccess) return -1;
    const int bv = max(1, min((a.n_vertices + kThreads - 1) / kThreads, 64));
    ROFTB_LAUNCH(k_render_vertices, dim3(bv, a.n_items), kThreads, 0, s, a, vt);
    const int per_block = kThreads / 32;
    const int bt = max(1, min((a.n_faces + per_block - 1) / per_block, max(1, 148 * 8 / a.n_items)));
    ROFTB_LAUNCH(k_render_triangles, dim3(bt, a.n_items), kThreads, 0, s, a, vt, zbuf);
    const int br = (int)max(1LL, min((npx + kThreads - 1) / kThreads, 148LL * 8));
    ROFTB_LAUNCH(k_render_resolve, br, kThreads, 0, s, zbuf, out, npx);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

size_t render_vertex_scratch_bytes(int n_items, int n_vertices) { return (size_t)n_items * n_vertices * sizeof(RVertex); }

int launch_pick_best(int n, const double* err, const int32_t* samples, double gain, int32_t* selected, double* likelihoods,
                     cudaStream_t s) {
    ROFTB_LAUNCH(k_pick_best, (n + 127) / 128, 128, 0, s, n, err, samples, gain, selected, likelihoods);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_or_models(int n_tracks, const int32_t* resume, const double* cand_mean, float* model, cudaStream_t s) {
    ROFTB_LAUNCH(k_or_models, (2 * n_tracks + 127) / 128, 128, 0, s, n_tracks, resume, cand_mean, model);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_or_select(int n_tracks, const int32_t* resume, const int32_t* selected, const double* cand_mean, const double* cand_cov,
                     double* mean, double* cov, cudaStream_t s) {
    ROFTB_LAUNCH(k_or_select, n_tracks, 64, 0, s, n_tracks, resume, selected, cand_mean, cand_cov, mean, cov);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace roftb
