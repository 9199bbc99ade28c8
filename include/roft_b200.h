/*
 * roft_b200.h - C ABI of libroft_b200.so: the B200-native (sm_100a) implementation of ROFT's
 * per-frame data-parallel hot path, batched over independent object tracks.
 *
 * Plain C, opaque handle, plain pointers and sizes; no exceptions cross this boundary.
 * Every entry point names the reference interface (file:line under hsp-iit/roft) it replaces.
 * The reference-side bindings a ROFT maintainer would add are shown in INTEGRATION.md; the
 * C++ adapter classes carrying the reference's names live in roft_b200/host/.
 *
 * Conventions
 *   - return value: 0 = ok, >0 = soft "no measurement / nothing done" (the reference's
 *     `return false` / pair<false,...>), <0 = hard error (bad argument, CUDA failure);
 *     roftb_last_error(ctx) gives the text of the last hard error.
 *   - one ctx per (GPU, host thread); calls on a ctx are serialised by the caller.
 *   - batch-first SoA: image planes are [n_tracks][H][W] row-major with a track stride in
 *     ELEMENTS; small per-track vectors are [n_tracks][k] FP64 in host memory.
 *   - quaternions are (w, x, y, z); the pose state vector is (v3, w3, x3, q4) = 13, its
 *     covariance 12x12 row-major (ROFTFilter.cpp:64-65,76-79).
 */
#ifndef ROFT_B200_H
#define ROFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ROFTB_VERSION 3

/* flow formats: the two cv::Mat types ROFT accepts (ImageOpticalFlowSource.h:44-45) */
#define ROFTB_FLOW_F32 13 /* CV_32FC2: float2 per element, grid 1, scale 1            */
#define ROFTB_FLOW_S16 11 /* CV_16SC2: short2 per element, NVOF1: grid 4, scale 32     */

#define ROFTB_MEM_HOST 0   /* pointers in roftb_frame are host memory (copied in, H2D)  */
#define ROFTB_MEM_DEVICE 1 /* pointers are device memory on ctx's GPU (zero copy)       */

/* pose measurement types (CartesianQuaternionMeasurement.h MeasurementType) */
#define ROFTB_MEAS_NONE 0
#define ROFTB_MEAS_VELOCITY 1
#define ROFTB_MEAS_POSE 2
#define ROFTB_MEAS_POSE_VELOCITY 3

#define ROFTB_MAX_DELAY 8 /* max frames between mask / pose iterations (D) in the filter loop */
#define ROFTB_MAX_CHAIN 30 /* longest flow chain of roftb_mask_sync (the stamped source's queue, ...Stamped.hpp:99) */

typedef struct roftb_ctx roftb_ctx;

/* Scalar parameters; roftb_config_default() fills config/config_fast_ycb.cfg values. */
typedef struct roftb_config {
    int32_t n_tracks;
    int32_t width, height;           /* camera_dataset.width/height (cfg:5-6)                      */
    double fx, fy, cx, cy;           /* cfg:7-10                                                   */
    double sample_time;              /* cfg:1                                                      */
    int32_t flow_format;             /* ROFTB_FLOW_*  (DatasetImageOpticalFlow.cpp:47)             */
    int32_t flow_grid;               /* width / flow.cols (DatasetImageOpticalFlow.cpp:46)         */
    float flow_scale;                /* 32 for CV_16SC2 else 1 (DatasetImageOpticalFlow.cpp:48-50) */
    /* measurement_model.velocity (cfg:79-85) */
    double cov_flow[2];
    double depth_maximum;
    int32_t subsampling_radius;
    int32_t weight_flow;
    /* kinematic_model.velocity / initial_condition.velocity (cfg:36-43,55-59) */
    double v_sigma[6];
    double v_cov0[6];
    /* kinematic_model.pose, initial_condition.pose, measurement_model.pose (cfg:23-34,49-53,71-77) */
    double p_sigma_linear[3];        /* psd of the linear acceleration                              */
    double p_sigma_angular[3];       /* variance of the angular velocity                            */
    double p_cov0[12];
    double cov_v[3], cov_w[3], cov_x[3], cov_q[3];
    double ut_alpha, ut_beta, ut_kappa; /* cfg:139-144 */
    int32_t use_pose, use_pose_resync, use_velocity, flow_aided; /* cfg:87-89,136 */
    int32_t segm_delay;              /* int(original_fps / desired_fps) of segmentation_dataset    */
    int32_t pose_delay;              /* same for pose_dataset                                      */
    int32_t device;                  /* CUDA device ordinal                                        */
    int32_t accum_fp64;              /* precision of the per-pixel Jacobian terms / normal-equation sums:
                                        1: FP64; 0: FP32 terms, FP64 reduction (agreement with the FP64 reference then
                                        scales as cond(Lambda)*6e-8/sqrt(N)); 2 (default): FP64 for tracks with fewer
                                        than 32768 candidate pixels, FP32 terms above */
    /* Render-and-compare pose outlier rejection inside the filter loop (outlier_rejection.enable / gain, cfg:108-112;
       ROFTFilter.cpp:346-359, 649-676).  Needs roftb_set_mesh before the first step.  The likelihood gain does not
       influence the choice (likelihood[0] > 2 likelihood[1]); the reference moreover receives it through a `const bool`
       constructor parameter (ROFTFilter.cpp:54), i.e. as 1.0 - pass what the caller's ROFTFilter would hold.
       divider: 0 = the reference's rule (2 for 640-wide frames, else 4: ROFTFilter.cpp:191-193). */
    int32_t outlier_rejection;
    int32_t outlier_rejection_divider;
    double outlier_rejection_gain;
} roftb_config;

/* One camera frame for all tracks = what the reference's sources deliver at one
 * ROFTFilter::filtering_step (ROFTFilter.cpp:255-367).  Image pointers are host or device
 * memory according to `memory`; the small per-track arrays are ALWAYS host memory.
 * With ROFTB_MEM_DEVICE the depth plane must stay valid and unmodified for 1 further step and
 * the flow plane for segm_delay further steps - ROFTB_MAX_DELAY further steps when segm_delay <= 0
 * (delay unknown: every flow since the last mask is chained) - (the reference clones them instead:
 * ImageOpticalFlowMeasurement.hpp:286, ImageSegmentationOFAidedSource.hpp:208).
 * Two buffers that are unbounded in the reference are bounded here: at most ROFTB_MAX_DELAY flows
 * are chained behind a new mask (with segm_delay <= 0 the oldest are dropped when more accumulate
 * between two masks), and at most ROFTB_MAX_DELAY + 3 buffered velocities are replayed by a pose
 * re-synchronisation (CartesianQuaternionMeasurement.cpp:100-104 trims to pose_delay + 1 anyway;
 * the cap only matters for pose_delay <= 0). */
typedef struct roftb_frame {
    int32_t memory;
    const float* depth;              /* [n_tracks][H][W] metres (CameraMeasurement.h:57)            */
    int64_t depth_track_stride;
    const void* flow;                /* [n_tracks][H/grid][W/grid][2]; NULL = no flow this frame    */
    int64_t flow_track_stride;       /* in scalar elements (float or int16)                         */
    const uint8_t* mask;             /* [n_tracks][H][W] (stale) mask delivered now; NULL = none    */
    int64_t mask_track_stride;
    const uint8_t* flow_valid;       /* host [n_tracks] or NULL (= valid wherever flow != NULL)     */
    const uint8_t* mask_valid;       /* host [n_tracks] or NULL (= valid wherever mask != NULL)     */
    const double* pose;              /* host [n_tracks][7] (x, q wxyz) or NULL                      */
    const uint8_t* pose_valid;       /* host [n_tracks] or NULL (= valid wherever pose != NULL)     */
    const double* dt;                /* host [n_tracks] elapsed camera time, or NULL = sample_time  */
} roftb_frame;

/* ---- context ------------------------------------------------------------------------- */
void roftb_config_default(roftb_config* cfg);
/* From 64 tracks on the context is built as two pipelined part contexts over the halves of the track range (DESIGN.md 6;
 * ROFTB_PARTS overrides) and owns one internal host thread; results do not depend on it.  A context is still to be used
 * from one host thread at a time. */
int roftb_create(const roftb_config* cfg, roftb_ctx** out);
void roftb_destroy(roftb_ctx* ctx);
const char* roftb_last_error(const roftb_ctx* ctx); /* ctx may be NULL: last create() error */
int roftb_sync(roftb_ctx* ctx);                     /* wait for all enqueued work           */
/* Non-blocking: make the main stream (roftb_stream) wait for everything enqueued so far on the library's
 * internal streams, so that an event recorded on roftb_stream afterwards covers the whole step. */
int roftb_join(roftb_ctx* ctx);
int roftb_version(void);
/* number of this library's kernel launches since create (bench.py "gpu_launches") */
int64_t roftb_kernel_launches(const roftb_ctx* ctx);
/* CUDA stream the kernels are enqueued on (cudaStream_t as void*), for event timing */
void* roftb_stream(roftb_ctx* ctx);
/* Device-side phase timing of roftb_filter_step with CUDA events.  Reads the averages (ms per step since
 * profiling was enabled) of 7 phases into ms_per_step[7] / steps (either may be NULL), then enables (1) or
 * disables (0) profiling; enabling resets the averages.  Phases: {preparation of the tracks with a new mask
 * (plane initialisation; prep stream), then the velocity kernel's device time split over its four stages in
 * proportion to per-track time stamps taken inside the kernel - worklist + pass A (gates, innovations, fused
 * mask propagation), pairing + median select, pass B (normal equations), kernel tail -, new-mask scatter
 * (own stream), pose UKF (own stream)}. */
int roftb_profile(roftb_ctx* ctx, int32_t enable, double* ms_per_step, int64_t* steps);

/* ---- the filter loop: ROFTFilter::initialization_step / filtering_step ----------------- */
/* ROFTFilter.cpp:216-237.  p_mean0: host [n_tracks][13] or NULL (zeros, q = identity);
 * v_mean0: host [n_tracks][6] or NULL. */
int roftb_filter_init(roftb_ctx* ctx, const double* p_mean0, const double* v_mean0);
/* ROFTFilter.cpp:255-367 for every track; asynchronous (returns once work is enqueued).
 * Argument errors (< 0) leave the context untouched; a CUDA failure in the middle of a step leaves the
 * per-track state machines ahead of the device work, so every later step fails until
 * roftb_filter_init() is called again. */
int roftb_filter_step(roftb_ctx* ctx, const roftb_frame* frame);
/* Blocking read-back of the beliefs (any pointer may be NULL): pose mean [T][13], pose
 * covariance [T][144], velocity mean [T][6], velocity covariance [T][36]. */
int roftb_get_state(roftb_ctx* ctx, double* p_mean, double* p_cov, double* v_mean, double* v_cov);
/* Blocking read-back of the synchronised mask of the last step: raw (the OF-aided source's
 * mask_, ImageSegmentationOFAidedSource.hpp:291-295) and/or thresholded (the
 * ImageSegmentationMeasurement output, ImageSegmentationMeasurement.cpp:61-65). Host [T][H][W]. */
int roftb_get_mask(roftb_ctx* ctx, uint8_t* raw, uint8_t* thresholded);
/* Diagnostics of the last velocity correction: number of valid flow pixels per track
 * (ImageOpticalFlowMeasurement.hpp:363-366), the accumulated information matrix
 * sum_j l_j H_j^T R^-1 H_j [T][36] and vector sum_j l_j H_j^T R^-1 z_j [T][6]. Blocking. */
int roftb_get_velocity_info(roftb_ctx* ctx, int32_t* count, double* lambda, double* eta);
/* Diagnostics of the last step's worklist: per track, the number of non-empty 128-pixel units of the
 * synchronised mask (units[T]) and the number of segmentation pixels in them (pixels[T]; the
 * findNonZero count of ImageOpticalFlowMeasurement.hpp:234 before subsampling and gates).  bench.py
 * derives the bytes the streaming passes had to touch from it.  Either pointer may be NULL. Blocking. */
int roftb_get_worklist(roftb_ctx* ctx, int32_t* units, int32_t* pixels);

/* ---- operators (stateless; buffers are HOST memory, copied through ctx scratch) --------- */
/* ImageSegmentationOFAidedSource<T>::map + cv::remap (ImageSegmentationOFAidedSource.hpp:
 * 211-226,235-281): warp `mask` through `n_flows` flow frames (oldest first).  zero_origin=1
 * reproduces the "no new mask" branch (mask_(0,0)=0 first, :224).  n_masks masks, each with
 * its own flow chain: flows is [n_flows][n_masks][H/grid][W/grid][2], n_flows <= ROFTB_MAX_CHAIN.
 * out_raw/out_thr: [n_masks][H][W].  The same call serves the time-stamped source
 * (ImageSegmentationOFAidedSourceStamped.hpp:153-318), whose host side only SELECTS the flows differently. */
int roftb_mask_sync(roftb_ctx* ctx, int32_t n_masks, const uint8_t* mask, const void* flows, int32_t n_flows,
                    int32_t zero_origin, uint8_t* out_raw, uint8_t* out_thr);
/* ImageOpticalFlowMeasurement<T>::freeze (hpp:231-283) fused with the Laplacian weights and
 * the sum of SKFCorrection.cpp:91-149 in information form, for n_items (mask, depth, flow)
 * triples: x_pred [n][6] is the predicted velocity used for the innovation norms; dt [n].
 * Outputs: lambda [n][36], eta [n][6], count [n] (valid pixels). */
int roftb_flow_velocity(roftb_ctx* ctx, int32_t n_items, const uint8_t* mask, const float* depth, const void* flow,
                        const double* x_pred, const double* dt, double* lambda, double* eta, int32_t* count);
/* SpatialVelocityModel + KFPrediction + SKFCorrection::correctStep (SpatialVelocityModel.cpp:
 * 15-27, SKFCorrection.cpp:37-153) in information form: x,P in/out [n][6],[n][36]. Applies the
 * observability gate of ROFTFilter.cpp:294-301 (count < 3 keeps the prior belief). */
int roftb_velocity_kf(roftb_ctx* ctx, int32_t n_items, const uint8_t* mask, const float* depth, const void* flow,
                      const double* dt, double* x, double* P, int32_t* count);
/* Materialised measurement for bfl-style callers: z [2N], H [2N][6] row-major, for ONE item
 * (ImageOpticalFlowMeasurement::measure / getMeasurementMatrix, hpp:297-326). capacity in pixels. */
int roftb_flow_measurement_export(roftb_ctx* ctx, const uint8_t* mask, const float* depth, const void* flow,
                                  double dt, int32_t capacity, double* z, double* H, int32_t* n_valid);
/* Masked depth de-projection (CameraMeasurement.cpp:75 PC mode restricted to the mask): points
 * [n][capacity][3] FP64 row-major order, count [n]. */
int roftb_masked_points(roftb_ctx* ctx, int32_t n_items, const uint8_t* mask, const float* depth, double max_depth,
                        int32_t capacity, double* points, int32_t* count);
/* Inner loop of ROFTFilter::pick_best_alternative (ROFTFilter.cpp:556-566): every 2nd mask pixel,
 * 0<d<2, rendered!=0: err_sum [n], samples [n]. rendered is [n][H/divider][W/divider]. */
int roftb_masked_depth_l1(roftb_ctx* ctx, int32_t n_items, const uint8_t* mask, const float* depth,
                          const float* rendered, int32_t divider, double* err_sum, int32_t* samples);
/* Mesh of the tracked object for the render-and-compare pose test: what ROFTFilter's constructor hands to SICAD
 * (ROFTFilter.cpp:184-199, MeshResource -> SICAD::ModelStreamContainer).  Assimp is not part of this path: the caller
 * passes the triangles (vertices [n_vertices][3] in the model frame, metres; faces [n_faces][3]).  Host memory. */
int roftb_set_mesh(roftb_ctx* ctx, const float* vertices, int32_t n_vertices, const int32_t* faces, int32_t n_faces);
/* Batched extension without a reference counterpart (one ROFTFilter tracks one object): per-track scale [n_tracks][3]
 * applied to the vertices of the shared mesh inside the filter loop's render-and-compare test, so that a batch can
 * hold objects of different sizes; NULL removes it.  Host memory. */
int roftb_set_mesh_scale(roftb_ctx* ctx, const float* scale);
/* SICAD::superimpose(poses, cam_x = 0, cam_o = identity, ..., depth) (SICAD.cpp:924-1066, depth attachment of
 * shader_model.frag:33-52) for a renderer built like ROFTFilter.cpp:194-198 (every intrinsic / divider, OpenGL-to-camera
 * rotation pi about x): poses [n_items][7] = (x, y, z, axis x y z, angle) as SICAD::ModelPose; out_depth
 * [n_items][H/divider][W/divider] = one tile per pose, metres, 0 where nothing is hit.  Host memory. */
int roftb_render_depth(roftb_ctx* ctx, int32_t n_items, const double* poses7, int32_t divider, float* out_depth);
/* ROFTFilter::pick_best_alternative (ROFTFilter.cpp:467-621) for n_items tracks: alternatives [n_items][2][13] are the
 * means of the two corrected beliefs (v, w, x, q wxyz), segmentation / depth [n_items][H][W] the (buffered) features;
 * renders both, takes the masked depth L1 of each (every 2nd mask pixel, 0 < d < 2, rendered != 0), likelihood = mean
 * error / gain (DBL_MAX without samples) and selects the second alternative iff likelihood[0] > 2 likelihood[1].
 * selected [n_items] (0 / 1), likelihoods [n_items][2] (may be NULL).  (The reference always reports success once the
 * renderer ran; a failed render is an error code here.) */
int roftb_pick_best_alternative(roftb_ctx* ctx, int32_t n_items, const uint8_t* segmentation, const float* depth,
                                const double* alternatives, int32_t divider, double gain, int32_t* selected,
                                double* likelihoods);
/* bfl::UKFPrediction through CartesianQuaternionModel (CartesianQuaternionModel.cpp:86-141):
 * mean [n][13], cov [n][144] in/out; dt [n]. */
int roftb_ukf_predict(roftb_ctx* ctx, int32_t n_items, double* mean, double* cov, const double* dt);
/* ROFT::UKFCorrection::correctStep through CartesianQuaternionMeasurement (UKFCorrection.cpp:
 * 54-133, CartesianQuaternionMeasurement.cpp:357-487): meas [n][13] laid out (v, w, x, q) with the
 * unused part ignored; meas_type [n] ROFTB_MEAS_*. */
int roftb_ukf_correct(roftb_ctx* ctx, int32_t n_items, double* mean, double* cov, const double* meas,
                      const int32_t* meas_type);

#ifdef __cplusplus
}
#endif
#endif /* ROFT_B200_H */
