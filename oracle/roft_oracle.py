"""CPU oracle for the ROFT per-frame hot path (TEST INFRASTRUCTURE - NOT PRODUCT CODE).

This module is a plain numpy (+ real OpenCV ``cv2`` primitives for the mask path)
restatement, in FP64, of the reference algorithm under ``/root/reference`` for the
hot path named by BASELINE.json.  It exists only as the checker for the CUDA
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
(``roft_b200``) never imports anything from ``oracle/``.

PARITY STATUS: **parity unpinned**.  The reference ships no golden vectors, no
known-answer tests and no assertions for this path and
it cannot be compiled here (Eigen3, OpenCV C++, BayesFilters, RobotsIO,
SuperimposeMesh, libconfig++, tclap are all absent; no network).  What *is*
pinned: the mask path uses the genuine OpenCV ``findNonZero`` / ``remap`` /
``threshold`` (cv2 4.13) exactly as the reference calls them, and every function
below cites the reference file:line it restates.  The dependency-free C++
restatement ``oracle/cpu_ref.cpp`` is cross-checked against this module.

Third-party arithmetic that is NOT under /root/reference (marked UPSTREAM-RECALL):
``robotology/bayes-filters-lib`` (unpinned: dockerfiles/Dockerfile:39-42 clones the
default branch) - UTWeight, sigma_point (SVD square root), unscented_transform,
sum_quaternion_rotation_vector, diff_quaternion, mean_quaternion,
GaussianMixture::augmentWithNoise, KFPrediction, UKFPrediction.  Their published
algorithm is restated from SURVEY.md Appendix B; two places where the published
description leaves the result under-determined are resolved here (and identically
in the CUDA path) and documented in DESIGN.md:
  * the quaternion log map takes the short way round (q and -q map to the same
    rotation vector), so the arbitrary sign of the eigenvector returned by
    ``mean_quaternion`` cannot change any covariance or innovation;
  * the mean quaternion is returned with a non-negative dot product with the
    first (central) sigma point, which fixes its sign deterministically.
"""
from __future__ import annotations

import math
from collections import deque
from dataclasses import dataclass, field
from typing import Deque, List, Optional, Sequence, Tuple

import numpy as np

try:  # genuine OpenCV primitives pin the bit-exact mask path
    import cv2
except Exception:  # pragma: no cover - cv2 is part of the image
    cv2 = None

INT_MIN = np.int32(-2147483648)
CV_32FC2 = 13
CV_16SC2 = 11


# --------------------------------------------------------------------------------------
# configuration (defaults = config/config_fast_ycb.cfg)
# --------------------------------------------------------------------------------------
@dataclass
class RoftConfig:
    """Scalar parameters of the path; defaults follow config/config_fast_ycb.cfg:1-144."""

    width: int = 1280
    height: int = 720
    fx: float = 1229.4285612615463
    fy: float = 1229.4285612615463
    cx: float = 640.0
    cy: float = 360.0
    sample_time: float = 0.033333333333  # cfg:1
    # flow source (DatasetImageOpticalFlow.cpp:46-50)
    flow_grid: int = 1
    flow_scale: float = 1.0
    # measurement_model.velocity (cfg:79-85)
    cov_flow: Tuple[float, float] = (1.0, 1.0)
    depth_maximum: float = 2.0
    subsampling_radius: float = 35.0
    weight_flow: bool = True
    # kinematic_model.velocity (cfg:55-59)
    v_sigma_linear: Tuple[float, float, float] = (0.1, 0.1, 0.1)
    v_sigma_angular: Tuple[float, float, float] = (0.1, 0.1, 0.1)
    # initial_condition.velocity (cfg:36-43)
    v_cov0: Tuple[float, ...] = (1e-3,) * 6
    # kinematic_model.pose (cfg:49-53): psd of linear acceleration, variance of angular velocity
    p_sigma_linear: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    p_sigma_angular: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    # initial_condition.pose (cfg:23-34)
    p_cov0: Tuple[float, ...] = (1e-3,) * 12
    # measurement_model.pose (cfg:71-77)
    cov_v: Tuple[float, float, float] = (0.1, 0.1, 0.1)
    cov_w: Tuple[float, float, float] = (1e-4, 1e-4, 1e-4)
    cov_x: Tuple[float, float, float] = (1e-3, 1e-3, 1e-3)
    cov_q: Tuple[float, float, float] = (1e-4, 1e-4, 1e-4)
    use_pose: bool = True
    use_pose_resync: bool = True
    use_velocity: bool = True
    flow_aided: bool = True
    # unscented_transform (cfg:139-144)
    ut_alpha: float = 1.0
    ut_beta: float = 2.0
    ut_kappa: float = 0.0
    # segmentation_dataset / pose_dataset (cfg:110-137): frames between iterations D
    segm_delay: int = 6
    pose_delay: int = 6
    # outlier_rejection (cfg:108-112).  The gain reaches ROFTFilter through a `const bool` parameter (ROFTFilter.cpp:54), so the
    # reference divides by 1.0; the choice does not depend on it.  divider: 0 = 2 for 640-wide frames else 4 (:191-193)
    outlier_rejection: bool = False
    outlier_rejection_gain: float = 1.0
    outlier_rejection_divider: int = 0


# --------------------------------------------------------------------------------------
# a1/a2  ImageOpticalFlowMeasurement
# --------------------------------------------------------------------------------------
def is_flow_valid(dx: np.ndarray, dy: np.ndarray) -> np.ndarray:
    """OpticalFlowUtilities.h:19-22 (float arguments)."""
    with np.errstate(invalid="ignore"):
        return (~np.isnan(dx)) & (~np.isnan(dy)) & (np.abs(dx) < 1e9) & (np.abs(dy) < 1e9)


def find_non_zero(mask: np.ndarray) -> np.ndarray:
    """cv::findNonZero: (x, y) int32 pairs in row-major order; empty -> shape (0, 2)."""
    if cv2 is not None:
        pts = cv2.findNonZero(np.ascontiguousarray(mask))
        if pts is None:
            return np.zeros((0, 2), np.int32)
        return pts.reshape(-1, 2).astype(np.int32)
    ys, xs = np.nonzero(mask)
    return np.stack([xs, ys], 1).astype(np.int32)


def _flow_at(flow: np.ndarray, rows: np.ndarray, cols: np.ndarray, scale: float) -> Tuple[np.ndarray, np.ndarray]:
    """float(f(k)) / flow_scaling_factor_ evaluated in FP32 (ImageOpticalFlowMeasurement.hpp:249-250)."""
    f = flow[rows, cols]
    s = np.float32(scale)
    dx = f[:, 0].astype(np.float32) / s
    dy = f[:, 1].astype(np.float32) / s
    return dx, dy


def flow_velocity_measurement(prev_mask: np.ndarray, prev_depth: np.ndarray, flow: np.ndarray,
                              cfg: RoftConfig, dt: float) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """ImageOpticalFlowMeasurement<T>::freeze, ImageOpticalFlowMeasurement.hpp:231-283.

    Returns (z [2N], H [2N,6], coords [N,2] as (u,v)) built from the PREVIOUS mask and
    depth and the CURRENT flow.  Selection is every ``subsampling_radius``-th non-zero
    pixel in row-major order (:237), gated AFTER selection (:252).
    """
    coords = find_non_zero(prev_mask)
    stride = int(np.float32(cfg.subsampling_radius))  # size_t ctor arg -> const float member (:99,:147)
    sel = coords[::stride]
    u = sel[:, 0].astype(np.int64)
    v = sel[:, 1].astype(np.int64)
    d = prev_depth[v, u].astype(np.float32)
    dx, dy = _flow_at(flow, v // cfg.flow_grid, u // cfg.flow_grid, cfg.flow_scale)
    ok = is_flow_valid(dx, dy) & (d > 0) & (d.astype(np.float64) < cfg.depth_maximum)
    u, v, d, dx, dy = u[ok], v[ok], d[ok].astype(np.float64), dx[ok], dy[ok]
    n = u.shape[0]
    z = np.empty(2 * n, np.float64)
    z[0::2] = dx.astype(np.float64)  # FP32 quotient widened (:272-273)
    z[1::2] = dy.astype(np.float64)
    uu = u.astype(np.float64) - cfg.cx
    vv = v.astype(np.float64) - cfg.cy
    fx, fy = cfg.fx, cfg.fy
    H = np.zeros((2 * n, 6), np.float64)
    H[0::2, 0] = fx / d
    H[0::2, 2] = -uu / d
    H[0::2, 3] = -uu * vv / fy
    H[0::2, 4] = fx + uu * uu / fx
    H[0::2, 5] = -vv * fx / fy
    H[1::2, 1] = fy / d
    H[1::2, 2] = -vv / d
    H[1::2, 3] = -(fy + vv * vv / fy)
    H[1::2, 4] = vv * uu / fx
    H[1::2, 5] = uu * fy / fx
    H *= dt  # measurement_matrix *= sample_time_ (:281)
    return z, H, np.stack([u, v], 1)


# --------------------------------------------------------------------------------------
# a3/a4  SKFCorrection + SpatialVelocityModel/KFPrediction
# --------------------------------------------------------------------------------------
def kf_predict(x: np.ndarray, P: np.ndarray, Q: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """bfl::KFPrediction with F = I6 (SpatialVelocityModel.cpp:15-27). UPSTREAM-RECALL."""
    return x.copy(), P + Q


def laplacian_stat_norms(innov: np.ndarray) -> np.ndarray:
    """The N "norms" whose median / mean absolute deviation fit the Laplacian (SKFCorrection.cpp:93-94).

    Reference quirk Q3 (replicated, not fixed): the reference views the interleaved innovation vector
    nu = [dx_0, dy_0, dx_1, dy_1, ...] (2N x 1) as ``Map<MatrixXd>(nu.data(), N, 2)``.  Eigen's MatrixXd is
    COLUMN-major, so row i of that map is (nu[i], nu[N + i]) - a component of valid pixel i // 2 paired with a
    component of valid pixel (N + i) // 2 - and ``rowwise().norm()`` gives r_i = sqrt(nu[i]^2 + nu[N+i]^2), NOT the
    per-pixel norm.  Only the likelihood loop (:111) uses the properly paired ``segment(2j, 2).norm()``.
    """
    n = innov.shape[0] // 2
    return np.sqrt(innov[:n] ** 2 + innov[n:2 * n] ** 2)


def laplacian_likelihoods(innov: np.ndarray) -> np.ndarray:
    """SKFCorrection.cpp:91-116: median / mean-absolute-deviation Laplacian re-weighting.

    ``innov`` is the interleaved 2N innovation vector of the valid pixels in selection order.  The median m and the
    scale b come from the column-major-mapped norms (Q3, ``laplacian_stat_norms``); the likelihood of pixel j is
    evaluated at its own norm ||(nu[2j], nu[2j+1])|| (:111) and the vector is divided by its true maximum (:114).
    """
    n = innov.shape[0] // 2
    s = np.sort(laplacian_stat_norms(innov))
    m = s[n // 2]
    if n % 2 == 0:
        m = 0.5 * (s[n // 2 - 1] + s[n // 2])
    b = np.abs(s - m).sum() / n
    lik = np.ones(n, np.float64)
    if b > 1e-4:
        pix = np.sqrt(innov[0::2] ** 2 + innov[1::2] ** 2)
        lik = np.maximum(1.0 / (2 * b) * np.exp(-np.abs(pix - m) / b), 1e-6)
        lik = lik / lik.max()
    return lik


def skf_correct(x_pred: np.ndarray, P_pred: np.ndarray, z: np.ndarray, H: np.ndarray,
                R: np.ndarray, weighting: bool) -> Tuple[np.ndarray, np.ndarray]:
    """SKFCorrection::correctStep, SKFCorrection.cpp:37-153: the SEQUENTIAL 2-row KF."""
    n = z.shape[0] // 2
    if n == 0:  # "measurement is empty" (:60-68)
        return x_pred.copy(), P_pred.copy()
    lik = None
    if weighting:
        innov = z - H @ x_pred  # innovations w.r.t. the predicted mean, computed once (:44,:74)
        lik = laplacian_likelihoods(innov)
    x = x_pred.astype(np.float64).copy()
    P = P_pred.astype(np.float64).copy()
    I6 = np.eye(6)
    for j in range(n):
        Hj = H[2 * j:2 * j + 2]
        Rj = R / lik[j] if weighting else R
        Py = Hj @ P @ Hj.T + Rj
        K = P @ Hj.T @ np.linalg.inv(Py)
        x = x + K @ (z[2 * j:2 * j + 2] - Hj @ x)
        P = (I6 - K @ Hj) @ P
    return x, P


def skf_correct_information(x_pred, P_pred, z, H, R, weighting):
    """Batch (information-form) equivalent of skf_correct - used as an identity test.

    Returns (x, P, Lambda_meas [6,6], eta_meas [6]) where Lambda_meas = sum l_j H_j^T R^-1 H_j.
    """
    n = z.shape[0] // 2
    if n == 0:
        return x_pred.copy(), P_pred.copy(), np.zeros((6, 6)), np.zeros(6)
    lik = np.ones(n)
    if weighting:
        innov = z - H @ x_pred
        lik = laplacian_likelihoods(innov)
    rinv = 1.0 / np.diag(R)
    w = np.empty(2 * n)
    w[0::2] = lik * rinv[0]
    w[1::2] = lik * rinv[1]
    Lm = H.T @ (H * w[:, None])
    em = H.T @ (w * z)
    Pinv = np.linalg.inv(P_pred)
    Lam = Pinv + Lm
    P = np.linalg.inv(Lam)
    x = P @ (Pinv @ x_pred + em)
    return x, P, Lm, em


# --------------------------------------------------------------------------------------
# a5-a8  delayed mask source, flow-aided mask synchronisation, threshold
# --------------------------------------------------------------------------------------
def _cvt_int(t: np.ndarray) -> np.ndarray:
    """C ``int(float)`` on x86 (cvttss2si): truncate toward zero; NaN/overflow -> INT_MIN."""
    t = np.asarray(t, np.float32)
    with np.errstate(invalid="ignore"):
        bad = ~np.isfinite(t) | (t >= np.float32(2147483648.0)) | (t < np.float32(-2147483648.0))
        out = np.where(bad, np.float32(0), np.trunc(t)).astype(np.int64)
    out[bad] = int(INT_MIN)
    return out


def mask_warp_map(mask: np.ndarray, flows: Sequence[np.ndarray], cfg: RoftConfig) -> np.ndarray:
    """ImageSegmentationOFAidedSource<T>::map, ImageSegmentationOFAidedSource.hpp:235-281.

    Returns the CV_32FC2 inverse map (zero-initialised, last writer wins in row-major
    source order).  Uses only the last ``segm_delay`` flows when segm_delay > 0 (:239-245).
    """
    H, W = mask.shape
    out = np.zeros((H, W, 2), np.float32)
    start = 0
    if cfg.segm_delay > 0:
        start = max(0, len(flows) - cfg.segm_delay)
    pts = find_non_zero(mask)
    if pts.shape[0] == 0:
        return out
    tx = pts[:, 0].astype(np.float32)
    ty = pts[:, 1].astype(np.float32)
    alive = np.ones(pts.shape[0], bool)
    g = np.float32(cfg.flow_grid)  # float / size_t -> float division (:269)
    s = np.float32(cfg.flow_scale)
    for j in range(start, len(flows)):
        ix, iy = _cvt_int(tx), _cvt_int(ty)
        inb = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
        alive &= inb  # error = true; break (:262-266)
        fr = _cvt_int(ty / g)
        fc = _cvt_int(tx / g)
        fr = np.where(alive, fr, 0)
        fc = np.where(alive, fc, 0)
        f = flows[j][fr, fc]
        with np.errstate(invalid="ignore", over="ignore"):
            ntx = (tx + f[:, 0].astype(np.float32) / s).astype(np.float32)
            nty = (ty + f[:, 1].astype(np.float32) / s).astype(np.float32)
        tx = np.where(alive, ntx, tx)
        ty = np.where(alive, nty, ty)
    ix, iy = _cvt_int(tx), _cvt_int(ty)
    ok = alive & (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
    # later sources overwrite earlier ones (:277): row-major order == ascending linear index
    src_lin = pts[:, 1].astype(np.int64) * W + pts[:, 0].astype(np.int64)
    winner = np.full(H * W, -1, np.int64)
    np.maximum.at(winner, (iy[ok] * W + ix[ok]), src_lin[ok])
    has = winner >= 0
    wy, wx = np.divmod(winner[has], W)
    flat = out.reshape(-1, 2)
    flat[has, 0] = wx.astype(np.float32)
    flat[has, 1] = wy.astype(np.float32)
    return out


def remap_exact(mask: np.ndarray, m: np.ndarray) -> np.ndarray:
    """cv::remap(mask, mask, map, Mat(), INTER_LINEAR, BORDER_CONSTANT) (hpp:215,225)."""
    if cv2 is not None:
        return cv2.remap(np.ascontiguousarray(mask), m, None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    sx = m[..., 0].astype(np.int64)
    sy = m[..., 1].astype(np.int64)
    return mask[sy, sx]


def threshold_mask(mask: np.ndarray) -> np.ndarray:
    """cv::threshold(seg, seg, 1, 255, THRESH_BINARY) (ImageSegmentationMeasurement.cpp:65)."""
    if cv2 is not None:
        return cv2.threshold(np.ascontiguousarray(mask), 1, 255, cv2.THRESH_BINARY)[1]
    return np.where(mask > 1, 255, 0).astype(np.uint8)


class DelayedMaskSchedule:
    """DatasetImageSegmentationDelayed::segmentation, DatasetImageSegmentationDelayed.cpp:42-63.

    ``index_for(head)`` returns the index of the (stale) frame delivered at head, or None.
    The same schedule is used by RobotsIO::DatasetTransformDelayed for poses (UPSTREAM-RECALL).
    """

    def __init__(self, delay: int, simulate_inference_time: bool = True, head0: int = 0):
        self.delay = int(delay)
        self.simulate = simulate_inference_time
        self.head0 = head0

    def index_for(self, head: int) -> Optional[int]:
        if self.delay <= 0:
            return head
        index = head - self.delay if self.simulate else head
        # C++ '%' truncates toward zero: remainder is 0 iff divisible
        if (index - self.head0) % self.delay != 0:
            return None
        if index < 0:
            index = self.head0
        return index


class OFAidedSegmentationSource:
    """ImageSegmentationOFAidedSource<T>::step_frame state machine, hpp:128-231."""

    def __init__(self, cfg: RoftConfig):
        self.cfg = cfg
        self.reset()

    def reset(self):  # hpp:299-314
        self.segmentation_available = False
        self.is_first_frame = True
        self.flow_buffer: List[np.ndarray] = []
        self.mask: Optional[np.ndarray] = None

    def step_frame(self, new_mask: Optional[np.ndarray], flow: Optional[np.ndarray]) -> bool:
        valid_segmentation = new_mask is not None
        mask = new_mask
        if (not self.segmentation_available) and valid_segmentation:  # :169-178
            self.segmentation_available = True
            self.mask = mask.copy()
            valid_segmentation = False
        if valid_segmentation:  # :180-197
            if find_non_zero(mask).shape[0] == 0:
                valid_segmentation = False
                if self.cfg.segm_delay <= 0:
                    self.flow_buffer.clear()
        valid_flow = (flow is not None) and (not self.is_first_frame)  # :200-204
        if valid_flow:
            self.flow_buffer.append(flow.copy())  # :208
        if valid_segmentation:  # :211-219
            self.mask = mask.copy()
            self.mask = remap_exact(self.mask, mask_warp_map(self.mask, self.flow_buffer, self.cfg))
            self.flow_buffer.clear()
        elif valid_flow and self.segmentation_available and self.mask is not None:  # :221-226
            self.mask = self.mask.copy()
            self.mask[0, 0] = 0
            self.mask = remap_exact(self.mask, mask_warp_map(self.mask, [flow], self.cfg))
        self.is_first_frame = False
        return True

    def segmentation(self) -> Tuple[bool, Optional[np.ndarray]]:  # :291-295
        return self.segmentation_available, self.mask


class OpticalFlowQueue:
    """OpticalFlowQueueHandler (src/OpticalFlowQueueHandler.cpp:18-57): the last ``window_size`` flow frames with the
    time stamp of the RGB frame they arrived with."""

    def __init__(self, window_size: int = 30):
        self.window_size = window_size
        self.buffer: Deque[Tuple[np.ndarray, float]] = deque()

    def add_flow(self, frame: np.ndarray, time_stamp: float):  # :18-26
        self.buffer.append((frame.copy(), float(time_stamp)))
        if len(self.buffer) > self.window_size:
            self.buffer.popleft()

    def get_buffer_region(self, initial_time_stamp: float) -> List[np.ndarray]:  # :29-51
        for index, (_, ts) in enumerate(self.buffer):
            if abs(ts - initial_time_stamp) < 1e-3:
                # the flow always refers to the PREVIOUS image: the region starts one entry later
                return [f for f, _ in list(self.buffer)[index + 1:]]
        return []

    def clear(self):
        self.buffer.clear()


class StampedOFAidedSegmentationSource:
    """ImageSegmentationOFAidedSourceStamped<T>::step_frame, ...Stamped.hpp:153-268: a new mask carries the time stamp
    of the frame it was computed on and is warped through the queued flows that FOLLOW that frame."""

    def __init__(self, cfg: RoftConfig):
        self.cfg = cfg
        self.segmentation_available = False
        self.is_first_frame = True
        self.queue = OpticalFlowQueue(30)  # flow_queue_max_size_ (:99)
        self.mask: Optional[np.ndarray] = None

    def step_frame(self, new_mask: Optional[np.ndarray], mask_time_stamp: float, flow: Optional[np.ndarray],
                   rgb_time_stamp: float) -> bool:
        valid_segmentation = new_mask is not None
        mask = new_mask
        if (not self.segmentation_available) and valid_segmentation:  # :212-221
            self.segmentation_available = True
            self.mask = mask.copy()
            valid_segmentation = False
        if valid_segmentation and find_non_zero(mask).shape[0] == 0:  # :223-230
            valid_segmentation = False
        valid_flow = (flow is not None) and (not self.is_first_frame)  # :233-236
        if valid_flow:
            self.queue.add_flow(flow, rgb_time_stamp)  # :237-241
        if valid_segmentation:  # :243-258
            self.mask = mask.copy()
            region = self.queue.get_buffer_region(mask_time_stamp)
            if len(region) > 0:
                self.mask = remap_exact(self.mask, mask_warp_map(self.mask, region, self.cfg))
            elif flow is not None:
                self.mask[0, 0] = 0
                self.mask = remap_exact(self.mask, mask_warp_map(self.mask, [flow], self.cfg))
        elif valid_flow and self.mask is not None:  # :260-265
            self.mask = self.mask.copy()
            self.mask[0, 0] = 0
            self.mask = remap_exact(self.mask, mask_warp_map(self.mask, [flow], self.cfg))
        self.is_first_frame = False
        return True

    def segmentation(self) -> Tuple[bool, Optional[np.ndarray]]:
        return self.segmentation_available, self.mask


# --------------------------------------------------------------------------------------
# a9  masked depth extraction
# --------------------------------------------------------------------------------------
def masked_points(mask: np.ndarray, depth: np.ndarray, cfg: RoftConfig, max_depth: float = 10.0) -> np.ndarray:
    """RobotsIO Camera::point_cloud restricted to the mask (CameraMeasurement.cpp:75; UPSTREAM-RECALL).

    X=(u-cx)d/fx, Y=(v-cy)d/fy, Z=d for mask pixels with 0<d<max_depth, row-major; returns [K,3] f64.
    """
    pts = find_non_zero(mask)
    u = pts[:, 0].astype(np.int64)
    v = pts[:, 1].astype(np.int64)
    d = depth[v, u].astype(np.float32)
    ok = (d > 0) & (d.astype(np.float64) < max_depth)
    u, v, d = u[ok], v[ok], d[ok].astype(np.float64)
    return np.stack([(u - cfg.cx) * d / cfg.fx, (v - cfg.cy) * d / cfg.fy, d], 1)


def masked_depth_l1(mask: np.ndarray, depth: np.ndarray, rendered: np.ndarray, divider: int) -> Tuple[float, int]:
    """Inner loop of ROFTFilter::pick_best_alternative, ROFTFilter.cpp:556-566.

    Every 2nd non-zero mask pixel; 0<d<2.0 and rendered!=0; returns (sum |d - render|, samples).
    """
    pts = find_non_zero(mask)[::2]
    u = pts[:, 0].astype(np.int64)
    v = pts[:, 1].astype(np.int64)
    d = depth[v, u].astype(np.float32)
    r = rendered[v // divider, u // divider].astype(np.float32)
    ok = (d > 0) & (d.astype(np.float64) < 2.0) & (r != 0)
    err = np.abs(d[ok] - r[ok]).astype(np.float64).sum()  # std::abs(float - float) accumulated in double
    return float(err), int(ok.sum())


# --------------------------------------------------------------------------------------
# bfl utilities (UPSTREAM-RECALL)
# --------------------------------------------------------------------------------------
def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Hamilton product, (w, x, y, z) ordering; broadcasts over leading dims."""
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw], -1)


def quat_conj(q: np.ndarray) -> np.ndarray:
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def rotation_vector_to_quaternion(r: np.ndarray) -> np.ndarray:
    """|r|>0 ? (cos(|r|/2), sin(|r|/2) r/|r|) : (1,0,0,0)."""
    r = np.atleast_2d(r)
    n = np.linalg.norm(r, axis=-1)
    q = np.zeros(r.shape[:-1] + (4,))
    q[..., 0] = 1.0
    nz = n > 0
    q[nz, 0] = np.cos(n[nz] / 2)
    q[nz, 1:] = np.sin(n[nz] / 2)[:, None] * r[nz] / n[nz][:, None]
    return q


def quaternion_to_rotation_vector(q: np.ndarray) -> np.ndarray:
    """log map: 2 acos(w) n/|n| taking the short way round (w<0 -> -q); |n|==0 -> 0."""
    q = np.atleast_2d(q).astype(np.float64)
    q = np.where(q[..., :1] < 0, -q, q)
    n = np.linalg.norm(q[..., 1:], axis=-1)
    out = np.zeros(q.shape[:-1] + (3,))
    nz = n > 0
    w = np.clip(q[nz, 0], -1.0, 1.0)
    out[nz] = (2.0 * np.arccos(w) / n[nz])[:, None] * q[nz, 1:]
    return out


def sum_quaternion_rotation_vector(q: np.ndarray, r: np.ndarray) -> np.ndarray:
    """exp_q(r) (x) q  (left multiplication)."""
    return quat_mul(rotation_vector_to_quaternion(r), np.asarray(q, np.float64))


def diff_quaternion(q_left: np.ndarray, q_right: np.ndarray) -> np.ndarray:
    """log_q(q_left (x) conj(q_right)) as a rotation vector."""
    return quaternion_to_rotation_vector(quat_mul(np.atleast_2d(q_left), quat_conj(np.asarray(q_right, np.float64))))


def mean_quaternion(weights: np.ndarray, quats: np.ndarray) -> np.ndarray:
    """Dominant eigenvector of sum_i w_i q_i q_i^T; sign fixed by the first column (see module doc)."""
    M = (quats.T * weights) @ quats
    M = 0.5 * (M + M.T)
    w, V = np.linalg.eigh(M)
    q = V[:, int(np.argmax(w))]
    if q @ quats[0] < 0:
        q = -q
    return q / np.linalg.norm(q)


def ut_weights(n: int, alpha: float, beta: float, kappa: float) -> Tuple[np.ndarray, np.ndarray, float]:
    """bfl::sigma_point::unscented_weights."""
    lam = alpha ** 2 * (n + kappa) - n
    wm = np.full(2 * n + 1, 1.0 / (2 * (n + lam)))
    wc = wm.copy()
    wm[0] = lam / (n + lam)
    wc[0] = lam / (n + lam) + (1 - alpha ** 2 + beta)
    return wm, wc, n + lam


JACOBI_REL_TOL = 1e-14
JACOBI_MAX_SWEEPS = 24


def jacobi_eigh(A: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Cyclic Jacobi eigen-decomposition of a symmetric PSD matrix (returns eigenvalues, V).

    A pair (p, q) is rotated only when |a_pq| > JACOBI_REL_TOL * sqrt(|a_pp a_qq|); the sweep
    loop ends when a whole sweep rotates nothing.  The relative test is what makes the result
    well defined on the exactly-degenerate covariances ROFT starts from (P = 1e-3 I: a diagonal
    block is left untouched, U = I there, as a two-sided Jacobi SVD does) - rounding noise in
    structurally-zero entries never triggers a rotation.  The CUDA path uses the same criterion
    (with a parallel pair ordering), see DESIGN.md.
    """
    A = np.array(A, np.float64)
    n = A.shape[0]
    V = np.eye(n)
    for _ in range(JACOBI_MAX_SWEEPS):
        rotated = False
        for p in range(n - 1):
            for q in range(p + 1, n):
                apq = A[p, q]
                if abs(apq) <= JACOBI_REL_TOL * math.sqrt(abs(A[p, p] * A[q, q])):
                    continue
                rotated = True
                theta = (A[q, q] - A[p, p]) / (2.0 * apq)
                t = math.copysign(1.0, theta) / (abs(theta) + math.sqrt(theta * theta + 1.0))
                c = 1.0 / math.sqrt(t * t + 1.0)
                s = t * c
                Ap = A[:, p].copy()
                Aq = A[:, q].copy()
                A[:, p] = c * Ap - s * Aq
                A[:, q] = s * Ap + c * Aq
                Ap = A[p, :].copy()
                Aq = A[q, :].copy()
                A[p, :] = c * Ap - s * Aq
                A[q, :] = s * Ap + c * Aq
                A[p, q] = 0.0
                A[q, p] = 0.0
                Vp = V[:, p].copy()
                Vq = V[:, q].copy()
                V[:, p] = c * Vp - s * Vq
                V[:, q] = s * Vp + c * Vq
        if not rotated:
            break
    return np.diag(A).copy(), V


def cov_sqrt(P: np.ndarray, method: str = "jacobi") -> np.ndarray:
    """A = U sqrt(S) with P = U S U^T (bfl::sigma_point::sigma_point uses a Jacobi SVD)."""
    if method == "svd":
        U, S, _ = np.linalg.svd(P)
        return U * np.sqrt(S)
    w, V = jacobi_eigh(0.5 * (P + P.T))
    return V * np.sqrt(np.maximum(w, 0.0))


def sigma_points(mean: np.ndarray, cov: np.ndarray, c: float, sqrt_method: str = "jacobi") -> np.ndarray:
    """bfl::sigma_point::sigma_point for a state (9 linear, 1 quaternion, k noise).

    mean: [13 + k]; cov: [12 + k, 12 + k]; returns [2n+1, 13 + k] (one sigma point per row).
    """
    n = cov.shape[0]
    A = cov_sqrt(cov, sqrt_method) * math.sqrt(c)
    pert = np.concatenate([np.zeros((1, n)), A.T, -A.T], 0)  # [2n+1, n]
    k = n - 12
    sp = np.empty((2 * n + 1, 13 + k))
    sp[:, :9] = mean[:9] + pert[:, :9]
    sp[:, 9:13] = sum_quaternion_rotation_vector(mean[9:13], pert[:, 9:12])
    sp[:, 13:] = mean[13:] + pert[:, 12:]
    return sp


# --------------------------------------------------------------------------------------
# a10  CartesianQuaternionModel
# --------------------------------------------------------------------------------------
def pose_model_Q(cfg: RoftConfig, T: float) -> np.ndarray:
    """CartesianQuaternionModel::evaluate_noise_covariance_matrix, CartesianQuaternionModel.cpp:127-141."""
    psd = np.diag(cfg.p_sigma_linear)
    Q = np.zeros((9, 9))
    Q[0:3, 0:3] = psd * T
    Q[3:6, 3:6] = np.diag(cfg.p_sigma_angular)
    Q[6:9, 6:9] = psd * (T ** 3.0 / 3.0)
    Q[0:3, 6:9] = psd * (T ** 2.0 / 2.0)
    Q[6:9, 0:3] = psd * (T ** 2.0 / 2.0)
    return Q


def pose_motion(sp: np.ndarray, T: float) -> np.ndarray:
    """CartesianQuaternionModel::motion, CartesianQuaternionModel.cpp:86-124 (rows = sigma points)."""
    out = np.empty((sp.shape[0], 13))
    out[:, :9] = sp[:, :9] + sp[:, 13:22]
    out[:, 6:9] += sp[:, 0:3] * T
    w = sp[:, 3:6]
    nw = np.linalg.norm(w, axis=1) + np.finfo(np.float64).eps
    q = sp[:, 9:13]
    dq = np.empty((sp.shape[0], 4))
    dq[:, 0] = np.cos(nw * T / 2.0)
    dq[:, 1:] = (np.sin(nw * T / 2.0) / nw)[:, None] * w
    out[:, 9:13] = quat_mul(dq, q)  # (cos I + sin/|w| Omega_left(w)) q
    return out


def ukf_predict(mean: np.ndarray, cov: np.ndarray, cfg: RoftConfig, T: float,
                sqrt_method: str = "jacobi") -> Tuple[np.ndarray, np.ndarray]:
    """bfl::UKFPrediction::predictStep through CartesianQuaternionModel (UPSTREAM-RECALL)."""
    Q = pose_model_Q(cfg, T)
    n = 21
    wm, wc, c = ut_weights(n, cfg.ut_alpha, cfg.ut_beta, cfg.ut_kappa)
    aug_mean = np.concatenate([mean, np.zeros(9)])
    aug_cov = np.zeros((21, 21))
    aug_cov[:12, :12] = cov
    aug_cov[12:, 12:] = Q
    sp = sigma_points(aug_mean, aug_cov, c, sqrt_method)
    prop = pose_motion(sp, T)
    out_mean = np.empty(13)
    out_mean[:9] = wm @ prop[:, :9]
    out_mean[9:] = mean_quaternion(wm, prop[:, 9:13])
    delta = np.empty((prop.shape[0], 12))
    delta[:, :9] = prop[:, :9] - out_mean[:9]
    delta[:, 9:] = diff_quaternion(prop[:, 9:13], out_mean[9:])
    out_cov = (delta.T * wc) @ delta
    return out_mean, out_cov


# --------------------------------------------------------------------------------------
# a11/a12  CartesianQuaternionMeasurement + ROFT::UKFCorrection
# --------------------------------------------------------------------------------------
MEAS_NONE, MEAS_VELOCITY, MEAS_POSE, MEAS_POSE_VELOCITY = 0, 1, 2, 3


def pose_meas_R(cfg: RoftConfig, mtype: int) -> np.ndarray:
    """CartesianQuaternionMeasurement.cpp:49-61."""
    Rv = np.diag(list(cfg.cov_v) + list(cfg.cov_w))
    Rp = np.diag(list(cfg.cov_x) + list(cfg.cov_q))
    if mtype == MEAS_VELOCITY:
        return Rv
    if mtype == MEAS_POSE:
        return Rp
    R = np.zeros((12, 12))
    R[:6, :6] = Rv
    R[6:, 6:] = Rp
    return R


def pose_predicted_measure(sp: np.ndarray, mtype: int) -> np.ndarray:
    """CartesianQuaternionMeasurement::predictedMeasure, .cpp:357-433 (use_screw_velocity == false)."""
    noise = sp[:, 13:]
    cols = []
    if mtype in (MEAS_POSE_VELOCITY, MEAS_VELOCITY):
        p = sp[:, 6:9]
        v = sp[:, 0:3]
        w = sp[:, 3:6]
        pv = v + np.cross(w, -p) + noise[:, 0:3]
        pw = w + noise[:, 3:6]
        cols += [pv, pw]
    if mtype in (MEAS_POSE_VELOCITY, MEAS_POSE):
        nx = noise[:, 0:3] if mtype == MEAS_POSE else noise[:, 6:9]
        px = sp[:, 6:9] + nx
        pq = quat_mul(rotation_vector_to_quaternion(noise[:, -3:]), sp[:, 9:13])
        cols += [px, pq]
    return np.concatenate(cols, 1)


def ukf_correct(mean: np.ndarray, cov: np.ndarray, meas: np.ndarray, mtype: int, cfg: RoftConfig,
                sqrt_method: str = "jacobi") -> Tuple[np.ndarray, np.ndarray]:
    """ROFT::UKFCorrection::correctStep, UKFCorrection.cpp:54-133.

    meas: [6] velocity, [7] pose (x, q wxyz) or [13] (v, w, x, q).
    """
    if mtype == MEAS_NONE:
        return mean.copy(), cov.copy()
    R = pose_meas_R(cfg, mtype)
    k = R.shape[0]
    n = 12 + k
    wm, wc, c = ut_weights(n, cfg.ut_alpha, cfg.ut_beta, cfg.ut_kappa)
    aug_mean = np.concatenate([mean, np.zeros(k)])
    aug_cov = np.zeros((n, n))
    aug_cov[:12, :12] = cov
    aug_cov[12:, 12:] = R
    sp = sigma_points(aug_mean, aug_cov, c, sqrt_method)
    y = pose_predicted_measure(sp, mtype)
    has_q = mtype in (MEAS_POSE_VELOCITY, MEAS_POSE)
    nlin = y.shape[1] - (4 if has_q else 0)
    ymean_lin = wm @ y[:, :nlin]
    dof = nlin + (3 if has_q else 0)
    dy = np.empty((y.shape[0], dof))
    dy[:, :nlin] = y[:, :nlin] - ymean_lin
    innovation = np.empty(dof)
    innovation[:nlin] = meas[:nlin] - ymean_lin
    if has_q:
        qm = mean_quaternion(wm, y[:, nlin:])
        dy[:, nlin:] = diff_quaternion(y[:, nlin:], qm)
        innovation[nlin:] = diff_quaternion(meas[nlin:], qm)[0]  # .cpp:456
    Py = (dy.T * wc) @ dy
    dx = np.empty((sp.shape[0], 12))
    dx[:, :9] = sp[:, :9] - mean[:9]
    dx[:, 9:] = diff_quaternion(sp[:, 9:13], mean[9:13])
    Pxy = (dx.T * wc) @ dy
    K = Pxy @ np.linalg.inv(Py)  # UKFCorrection.cpp:118
    Kn = K @ innovation
    out = np.empty(13)
    out[:9] = mean[:9] + Kn[:9]
    out[9:] = sum_quaternion_rotation_vector(mean[9:13], Kn[9:12])[0]
    return out, cov - K @ Py @ K.T


class PoseMeasurementModel:
    """CartesianQuaternionMeasurement::freeze mode machine, .cpp:92-348 (use_screw_velocity=false)."""

    STANDARD, POP_BUFFERED, REPEAT_ONLY_VELOCITY = 0, 1, 2

    def __init__(self, cfg: RoftConfig):
        self.cfg = cfg
        self.buffer: Deque[np.ndarray] = deque()
        self.is_pose = False
        self.is_first_velocity_in = False
        self.last_v = np.zeros(3)
        self.last_w = np.zeros(3)
        self.last_pose: Optional[np.ndarray] = None  # (x, q wxyz)
        self.mtype = MEAS_NONE
        self.measurement = np.zeros(0)

    def _set(self, mtype):
        self.mtype = mtype
        vel = np.concatenate([self.last_v, self.last_w])
        if mtype == MEAS_POSE_VELOCITY:
            self.measurement = np.concatenate([vel, self.last_pose])
        elif mtype == MEAS_VELOCITY:
            self.measurement = vel
        elif mtype == MEAS_POSE:
            self.measurement = self.last_pose.copy()

    def freeze(self, mode: int, velocity: Optional[np.ndarray] = None, pose: Optional[np.ndarray] = None) -> bool:
        cfg = self.cfg
        if mode == self.POP_BUFFERED:  # :97-152
            if cfg.pose_delay > 0:
                while len(self.buffer) > cfg.pose_delay + 1:
                    self.buffer.popleft()
            if len(self.buffer) == 0:
                self.buffer.append(self.measurement[:6].copy())
                return False
            bv = self.buffer.popleft()
            self.last_v, self.last_w = bv[:3].copy(), bv[3:].copy()
            if self.is_pose:
                self._set(MEAS_POSE_VELOCITY)
                self.is_pose = False
            else:
                self._set(MEAS_VELOCITY)
            return True
        if mode == self.REPEAT_ONLY_VELOCITY:  # :154-174
            if self.is_first_velocity_in:
                self._set(MEAS_VELOCITY)
            return True
        # Standard (:176-347)
        if cfg.use_velocity and velocity is not None:
            self.is_first_velocity_in = True
            self.last_v, self.last_w = velocity[:3].copy(), velocity[3:].copy()
        self.is_pose = False
        if cfg.use_pose:
            self.is_pose = pose is not None
            if self.is_pose:
                self.last_pose = pose.copy()
        if self.is_first_velocity_in and self.is_pose:
            self._set(MEAS_POSE_VELOCITY)
            self.buffer.append(self.measurement[:6].copy())
        elif self.is_first_velocity_in:
            self._set(MEAS_VELOCITY)
            self.buffer.append(self.measurement.copy())
        elif self.is_pose:
            self._set(MEAS_POSE)
        else:
            self.mtype = MEAS_NONE
            return False
        return True


# --------------------------------------------------------------------------------------
# a13  ROFTFilter::filtering_step (without the OpenGL render-and-compare)
# --------------------------------------------------------------------------------------
@dataclass
class FrameInput:
    """What the reference's sources deliver at one frame (all optional but depth)."""
    depth: np.ndarray                       # float32 [H, W]
    flow: Optional[np.ndarray] = None       # [Hf, Wf, 2] float32 or int16; None = not available
    mask: Optional[np.ndarray] = None       # uint8 [H, W]: a (stale) mask delivered this frame
    pose: Optional[np.ndarray] = None       # [7] x, q(wxyz): a (stale) pose delivered this frame
    dt: Optional[float] = None              # elapsed camera time; None -> sample_time


class RoftFilterOracle:
    """Single-track restatement of ROFTFilter (ROFTFilter.cpp:216-367) over the functions above.

    With cfg.outlier_rejection the render-and-compare pose test (ROFTFilter.cpp:467-621, 649-676) runs over the
    restated rasteriser below (`mesh` = (vertices, faces) of the tracked object); everything in filtering_step is
    followed in order.
    """

    def __init__(self, cfg: RoftConfig, x0: Optional[np.ndarray] = None, sqrt_method: str = "jacobi", mesh=None):
        self.cfg = cfg
        self.mesh = mesh
        self.or_features = None          # (segmentation, depth) of buffer_outlier_rejection_features
        self.or_selected: List[int] = []  # diagnostics: choices made so far
        self.sqrt_method = sqrt_method
        self.v_mean = np.zeros(6)
        self.v_cov = np.diag(cfg.v_cov0).astype(np.float64)
        self.p_mean = np.zeros(13)
        self.p_mean[9] = 1.0
        if x0 is not None:
            self.p_mean[:] = x0
        self.p_cov = np.diag(cfg.p_cov0).astype(np.float64)
        self.buffered = (self.p_mean.copy(), self.p_cov.copy())
        self.Qv = np.diag(list(cfg.v_sigma_linear) + list(cfg.v_sigma_angular))
        self.Rflow = np.diag(cfg.cov_flow)
        self.seg_source = OFAidedSegmentationSource(cfg)
        self.pose_model = PoseMeasurementModel(cfg)
        # ImageSegmentationMeasurement state
        self.seg_available = False
        self.seg: Optional[np.ndarray] = None
        # ImageOpticalFlowMeasurement state
        self.flow_first_frame = True
        self.prev_depth: Optional[np.ndarray] = None
        self.prev_seg: Optional[np.ndarray] = None
        # diagnostics of the last step
        self.last_n_valid = 0
        self.last_info = None

    def step(self, fr: FrameInput):
        cfg = self.cfg
        dt = cfg.sample_time if fr.dt is None else fr.dt
        # 4. segmentation_->freeze() (ImageSegmentationMeasurement.cpp:30-75)
        if cfg.flow_aided:
            self.seg_source.step_frame(fr.mask, fr.flow)
            new_seg, raw = self.seg_source.segmentation()
        else:
            new_seg, raw = (fr.mask is not None), fr.mask
        if new_seg:
            self.seg_available = True
            self.seg = threshold_mask(raw.copy())
        data_in = self.seg_available
        # 5. flow measurement freeze (ImageOpticalFlowMeasurement.hpp:184-293)
        meas = None
        if self.seg_available:
            if fr.flow is None or self.flow_first_frame:
                self.prev_depth, self.prev_seg = fr.depth, self.seg
                self.flow_first_frame = False
                data_in = False
            else:
                meas = flow_velocity_measurement(self.prev_seg, self.prev_depth, fr.flow, cfg, dt)
                self.prev_depth, self.prev_seg = fr.depth, self.seg
        self.last_n_valid = 0
        self.last_info = None
        # 6. velocity KF (ROFTFilter.cpp:291-302)
        if data_in:
            z, H, _ = meas
            xp, Pp = kf_predict(self.v_mean, self.v_cov, self.Qv)
            xc, Pc = skf_correct(xp, Pp, z, H, self.Rflow, cfg.weight_flow)
            self.last_n_valid = z.shape[0] // 2
            self.last_info = (xp, Pp, z, H)
            if z.shape[0] // 2 >= 3:  # check_observability (hpp:363-366)
                self.v_mean, self.v_cov = xc, Pc
        # 7-9. pose UKF (ROFTFilter.cpp:305-367)
        velocity = self.v_mean.copy()  # velocity_->set_twist(...) then freeze(true) always succeeds
        T = dt
        # ROFTFilter.cpp:313-321: the first step of a re-synchronising filter buffers the features of the test
        current_features = (self.seg if self.seg is not None else np.zeros((cfg.height, cfg.width), np.uint8), fr.depth)
        if cfg.outlier_rejection and cfg.use_pose_resync and self.or_features is None:
            self.or_features = current_features
        pm, pc = ukf_predict(self.p_mean, self.p_cov, cfg, T, self.sqrt_method)
        pmodel = self.pose_model
        if pmodel.freeze(pmodel.STANDARD, velocity, fr.pose):
            if pmodel.mtype == MEAS_POSE_VELOCITY and cfg.use_pose_resync:
                buffered_copy = self.buffered
                self.buffered = (self.p_mean.copy(), self.p_cov.copy())
                cm, cc = buffered_copy
                while pmodel.freeze(pmodel.POP_BUFFERED):
                    pm, pc = ukf_predict(cm, cc, cfg, T, self.sqrt_method)
                    if cfg.outlier_rejection and pmodel.mtype == MEAS_POSE_VELOCITY:
                        cm, cc = self._correct_outlier_rejection(pm, pc, self.or_features)
                    else:
                        cm, cc = ukf_correct(pm, pc, pmodel.measurement, pmodel.mtype, cfg, self.sqrt_method)
                self.p_mean, self.p_cov = cm, cc
                if cfg.outlier_rejection:
                    self.or_features = current_features  # :352-353
            elif cfg.outlier_rejection and pmodel.mtype == MEAS_POSE_VELOCITY:
                self.p_mean, self.p_cov = self._correct_outlier_rejection(pm, pc, current_features)  # :357-358
            else:
                self.p_mean, self.p_cov = ukf_correct(pm, pc, pmodel.measurement, pmodel.mtype, cfg, self.sqrt_method)
        else:
            self.p_mean, self.p_cov = pm, pc
        return self.p_mean.copy(), self.v_mean.copy()

    def _correct_outlier_rejection(self, pm, pc, features):
        """ROFTFilter::correct_outlier_rejection, ROFTFilter.cpp:649-676."""
        cfg, pmodel = self.cfg, self.pose_model
        a_m, a_c = ukf_correct(pm, pc, pmodel.measurement, pmodel.mtype, cfg, self.sqrt_method)
        pmodel.freeze(pmodel.REPEAT_ONLY_VELOCITY)
        b_m, b_c = ukf_correct(pm, pc, pmodel.measurement, pmodel.mtype, cfg, self.sqrt_method)
        seg, depth = features
        divider = cfg.outlier_rejection_divider or (2 if cfg.width == 640 else 4)
        sel, _ = pick_best_alternative(cfg, self.mesh[0], self.mesh[1], [a_m, b_m], seg, depth, divider, cfg.outlier_rejection_gain)
        self.or_selected.append(sel)
        return (b_m, b_c) if sel == 1 else (a_m, a_c)


# ---- pose outlier rejection ------------------------------------------------------------------
# PARITY UNPINNED: the reference renders with OpenGL (SICAD) whose rasterisation is implementation-defined in the last
# bits and cannot run here (no GL stack).  The functions below restate the pipeline the reference drives it with; the
# CUDA rasteriser (roft_b200/csrc/render.cu) follows the same restatement and is compared against it with a tolerance.

def quaternion_to_axis_angle(q: np.ndarray) -> np.ndarray:
    """Eigen::AngleAxisd(Eigen::Quaterniond(w, x, y, z)) (ROFTFilter.cpp:518-524; UPSTREAM-RECALL of Eigen 3.3+)."""
    w, v = float(q[0]), np.asarray(q[1:4], np.float64)
    n = float(np.sqrt(v @ v))
    if n != 0.0:
        angle = 2.0 * np.arctan2(n, abs(w))
        if w < 0.0:
            n = -n
        return np.array([v[0] / n, v[1] / n, v[2] / n, angle])
    return np.array([1.0, 0.0, 0.0, 0.0])


def sicad_model_matrix(pose7: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """glm::rotate(I, float(angle), float(axis)) + translation (SICAD.cpp:604-607; UPSTREAM-RECALL of glm), float32."""
    f = np.float32
    ang = f(pose7[6])
    ax = np.asarray(pose7[3:6], np.float64).astype(f)
    c, s = f(np.cos(ang)), f(np.sin(ang))
    n = f(np.sqrt(f(f(ax[0] * ax[0] + ax[1] * ax[1]) + ax[2] * ax[2])))
    if n > 0:
        ax = (ax / n).astype(f)
    t = ((f(1) - c) * ax).astype(f)
    x, y, z = ax
    R = np.array([[c + t[0] * x, t[1] * x - s * z, t[2] * x + s * y],
                  [t[0] * y + s * z, c + t[1] * y, t[2] * y - s * x],
                  [t[0] * z - s * y, t[1] * z + s * x, c + t[2] * z]], f)
    return R, np.asarray(pose7[:3], np.float64).astype(f)


def render_depth(vertices: np.ndarray, faces: np.ndarray, pose7: np.ndarray, width: int, height: int,
                 fx: float, fy: float, cx: float, cy: float) -> np.ndarray:
    """SICAD::superimpose depth output for one pose (SICAD.cpp:924-1066, shader_model.frag:33-52).

    Projection matrix of SICAD.cpp:1634-1637 + viewport + cv::flip reduce to u = fx X/Z + cx, v = fy Y/Z + cy with pixel
    centres at +0.5; window depth (z_ndc + 1) / 2 for near 0.001 / far 1000 in float32; vertices snapped to 1/256 px;
    inclusive integer edge functions; z interpolated with the integer barycentrics in float64, rounded to float32;
    nearest fragment wins; linearize_depth() in float32; 0 where nothing is hit.  No clipping: triangles with a
    vertex outside (near, far) are dropped.
    """
    f = np.float32
    near, far = f(0.001), f(1000.0)
    R, t = sicad_model_matrix(pose7)
    V = np.asarray(vertices, f)
    P = np.empty_like(V)
    for r in range(3):
        P[:, r] = ((R[r, 0] * V[:, 0] + R[r, 1] * V[:, 1]).astype(f) + R[r, 2] * V[:, 2]).astype(f) + t[r]
    X, Y, Z = P[:, 0], P[:, 1], P[:, 2]
    ok = (Z > near) & (Z < far)
    iz = (f(1) / np.where(ok, Z, f(1))).astype(f)
    u = ((f(fx) * X).astype(f) * iz).astype(f) + f(cx)
    v = ((f(fy) * Y).astype(f) * iz).astype(f) + f(cy)
    A = f((far + near) / (far - near))
    B = f(f(f(2) * far * near) / (far - near))
    zn = (A - (B * iz).astype(f)).astype(f)
    zw = (f(0.5) * zn).astype(f) + f(0.5)
    lim = f(1.0e6 * 256)
    xs = np.rint(np.clip((u * f(256)).astype(f), -lim, lim)).astype(np.int64)
    ys = np.rint(np.clip((v * f(256)).astype(f), -lim, lim)).astype(np.int64)
    zbuf = np.full((height, width), np.inf, f)
    for tri in np.asarray(faces, np.int64):
        i0, i1, i2 = tri
        if not (ok[i0] and ok[i1] and ok[i2]):
            continue
        area = (xs[i1] - xs[i0]) * (ys[i2] - ys[i0]) - (ys[i1] - ys[i0]) * (xs[i2] - xs[i0])
        if area == 0:
            continue
        if area < 0:
            i1, i2, area = i2, i1, -area
        x0, x1, x2, y0, y1, y2 = xs[i0], xs[i1], xs[i2], ys[i0], ys[i1], ys[i2]
        xmin = max(0, (min(x0, x1, x2) - 128 + 255) >> 8)
        xmax = min(width - 1, (max(x0, x1, x2) - 128) >> 8)
        ymin = max(0, (min(y0, y1, y2) - 128 + 255) >> 8)
        ymax = min(height - 1, (max(y0, y1, y2) - 128) >> 8)
        if xmin > xmax or ymin > ymax:
            continue
        px = np.arange(xmin, xmax + 1, dtype=np.int64)[None, :] * 256 + 128
        py = np.arange(ymin, ymax + 1, dtype=np.int64)[:, None] * 256 + 128
        w0 = (x2 - x1) * (py - y1) - (y2 - y1) * (px - x1)
        w1 = (x0 - x2) * (py - y2) - (y0 - y2) * (px - x2)
        w2 = area - w0 - w1
        inside = (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
        zd = (w0.astype(np.float64) * float(zw[i0]) + w1.astype(np.float64) * float(zw[i1]) +
              w2.astype(np.float64) * float(zw[i2])) * (1.0 / float(area))
        zf = zd.astype(f)
        inside &= (zf >= 0) & (zf <= 1)
        sub = zbuf[ymin:ymax + 1, xmin:xmax + 1]
        np.minimum(sub, np.where(inside, zf, np.inf).astype(f), out=sub)
    hit = np.isfinite(zbuf)
    z = (np.where(hit, zbuf, f(0)) * f(2)).astype(f) - f(1)
    lin = (f(f(2) * near * far) / ((far + near) - (z * (far - near)).astype(f)).astype(f)).astype(f)
    return np.where(hit, lin, f(0)).astype(f)


def pick_best_alternative(cfg: "RoftConfig", vertices: np.ndarray, faces: np.ndarray, alternatives: Sequence[np.ndarray],
                          segmentation: np.ndarray, depth: np.ndarray, divider: int, gain: float):
    """ROFTFilter::pick_best_alternative, ROFTFilter.cpp:467-621: (selected, likelihoods)."""
    w, h = cfg.width // divider, cfg.height // divider
    lik = []
    for alt in alternatives:
        aa = quaternion_to_axis_angle(np.asarray(alt[9:13]))
        pose7 = np.concatenate([np.asarray(alt[6:9], np.float64), aa])
        rend = render_depth(vertices, faces, pose7, w, h, cfg.fx / divider, cfg.fy / divider, cfg.cx / divider, cfg.cy / divider)
        err, n = masked_depth_l1(segmentation, depth, rend, divider)
        lik.append(np.finfo(np.float64).max if n == 0 else (err / n) / gain)
    with np.errstate(over="ignore"):  # 2 * DBL_MAX = inf, as in C
        selected = 1 if lik[0] > np.float64(2.0) * np.float64(lik[1]) else 0
    return selected, np.array(lik)
