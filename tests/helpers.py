"""Shared helpers for the parity tests (test infrastructure; may import the oracle)."""
from __future__ import annotations

import os

import numpy as np

import roft_oracle as o
from roft_b200.synthetic import make_sequence


def rel(a, b, eps=1e-12):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), eps))


def small_cfg(W=320, H=180, **kw):
    base = dict(width=W, height=H, fx=1229.4285612615463 * W / 1280, fy=1229.4285612615463 * W / 1280,
                cx=W / 2.0, cy=H / 2.0)
    base.update(kw)
    return o.RoftConfig(**base)


def to_roftb_config(cfg: o.RoftConfig, n_tracks: int, flow_format="f32"):
    from roft_b200 import api
    return api.default_config(
        n_tracks=n_tracks, width=cfg.width, height=cfg.height, fx=cfg.fx, fy=cfg.fy, cx=cfg.cx, cy=cfg.cy,
        sample_time=cfg.sample_time, flow_format=api.FLOW_F32 if flow_format == "f32" else api.FLOW_S16,
        flow_grid=cfg.flow_grid, flow_scale=cfg.flow_scale, cov_flow=cfg.cov_flow, depth_maximum=cfg.depth_maximum,
        subsampling_radius=int(cfg.subsampling_radius), weight_flow=int(cfg.weight_flow),
        v_sigma=tuple(cfg.v_sigma_linear) + tuple(cfg.v_sigma_angular), v_cov0=cfg.v_cov0,
        p_sigma_linear=cfg.p_sigma_linear, p_sigma_angular=cfg.p_sigma_angular, p_cov0=cfg.p_cov0,
        cov_v=cfg.cov_v, cov_w=cfg.cov_w, cov_x=cfg.cov_x, cov_q=cfg.cov_q,
        ut_alpha=cfg.ut_alpha, ut_beta=cfg.ut_beta, ut_kappa=cfg.ut_kappa,
        use_pose=int(cfg.use_pose), use_pose_resync=int(cfg.use_pose_resync), use_velocity=int(cfg.use_velocity),
        flow_aided=int(cfg.flow_aided), segm_delay=cfg.segm_delay, pose_delay=cfg.pose_delay,
        outlier_rejection=int(cfg.outlier_rejection), outlier_rejection_gain=float(cfg.outlier_rejection_gain),
        outlier_rejection_divider=int(cfg.outlier_rejection_divider),
        accum_fp64=int(os.environ.get("ROFTB_TEST_ACCUM_FP64", "2")))


def sequence(cfg: o.RoftConfig, n_tracks, n_frames, seed=3, flow_format="f32", **kw):
    return make_sequence(n_tracks, n_frames, cfg.width, cfg.height, cfg.fx, cfg.fy, cfg.cx, cfg.cy,
                         dt=cfg.sample_time, seed=seed, flow_format=flow_format, **kw)


def frame_inputs(seq, cfg: o.RoftConfig, k: int, t: int):
    """What the delayed dataset sources deliver to track t at frame k (DatasetImageSegmentationDelayed.cpp:42-63)."""
    ms = o.DelayedMaskSchedule(cfg.segm_delay).index_for(k)
    ps = o.DelayedMaskSchedule(cfg.pose_delay).index_for(k)
    mask = seq.mask[ms, t].numpy() if ms is not None else None
    pose = seq.pose[ps, t].numpy() if (ps is not None and bool(seq.pose_valid[ps, t])) else None
    flow = seq.flow[k, t].numpy() if k > 0 else None
    return o.FrameInput(depth=seq.depth[k, t].numpy(), flow=flow, mask=mask, pose=pose, dt=cfg.sample_time)


def quat_close(a, b):
    """distance between unit quaternions up to sign"""
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b))
