"""ctypes wrapper of oracle/libcpu_ref.so (TEST INFRASTRUCTURE / REPORTED CPU BASELINE - not product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class CrefConfig(C.Structure):
    _fields_ = [
        ("W", C.c_int32), ("H", C.c_int32),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("sample_time", C.c_double),
        ("flow_s16", C.c_int32), ("grid", C.c_int32), ("scale", C.c_float),
        ("cov_flow", C.c_double * 2), ("depth_max", C.c_double),
        ("stride", C.c_int32), ("weight_flow", C.c_int32),
        ("v_sigma", C.c_double * 6), ("v_cov0", C.c_double * 6),
        ("psd_lin", C.c_double * 3), ("sigma_ang", C.c_double * 3), ("p_cov0", C.c_double * 12),
        ("cov_v", C.c_double * 3), ("cov_w", C.c_double * 3), ("cov_x", C.c_double * 3), ("cov_q", C.c_double * 3),
        ("alpha", C.c_double), ("beta", C.c_double), ("kappa", C.c_double),
        ("use_pose", C.c_int32), ("use_pose_resync", C.c_int32), ("use_velocity", C.c_int32), ("flow_aided", C.c_int32),
        ("segm_delay", C.c_int32), ("pose_delay", C.c_int32),
    ]


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_DIR, "libcpu_ref.so")
        src = os.path.join(_DIR, "cpu_ref.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", _DIR])
        _LIB = C.CDLL(so)
        _LIB.cref_filter_create.restype = C.c_void_p
        _LIB.cref_filter_create.argtypes = [C.POINTER(CrefConfig), C.c_void_p]
        _LIB.cref_filter_destroy.argtypes = [C.c_void_p]
        _LIB.cref_filter_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        _LIB.cref_filter_state.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        _LIB.cref_filter_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.cref_flow_measurement.restype = C.c_int32
        _LIB.cref_flow_measurement.argtypes = [C.POINTER(CrefConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32,
                                               C.c_void_p, C.c_void_p]
        _LIB.cref_skf_correct.argtypes = [C.POINTER(CrefConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        _LIB.cref_mask_warp.argtypes = [C.POINTER(CrefConfig), C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]
        _LIB.cref_ukf_predict.argtypes = [C.POINTER(CrefConfig), C.c_void_p, C.c_void_p, C.c_double]
        _LIB.cref_ukf_correct.argtypes = [C.POINTER(CrefConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        _LIB.cref_timed_run.restype = C.c_int64
        _LIB.cref_timed_run.argtypes = [C.POINTER(CrefConfig), C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


def make_config(cfg) -> CrefConfig:
    """cfg: roft_oracle.RoftConfig"""
    c = CrefConfig()
    c.W, c.H = cfg.width, cfg.height
    c.fx, c.fy, c.cx, c.cy, c.sample_time = cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.sample_time
    c.grid, c.scale = cfg.flow_grid, cfg.flow_scale
    c.flow_s16 = 1 if cfg.flow_scale != 1.0 else 0
    c.cov_flow[0], c.cov_flow[1] = cfg.cov_flow
    c.depth_max = cfg.depth_maximum
    c.stride = int(cfg.subsampling_radius)
    c.weight_flow = int(cfg.weight_flow)
    for i in range(3):
        c.v_sigma[i] = cfg.v_sigma_linear[i]; c.v_sigma[3 + i] = cfg.v_sigma_angular[i]
        c.psd_lin[i] = cfg.p_sigma_linear[i]; c.sigma_ang[i] = cfg.p_sigma_angular[i]
        c.cov_v[i] = cfg.cov_v[i]; c.cov_w[i] = cfg.cov_w[i]; c.cov_x[i] = cfg.cov_x[i]; c.cov_q[i] = cfg.cov_q[i]
    for i in range(6):
        c.v_cov0[i] = cfg.v_cov0[i]
    for i in range(12):
        c.p_cov0[i] = cfg.p_cov0[i]
    c.alpha, c.beta, c.kappa = cfg.ut_alpha, cfg.ut_beta, cfg.ut_kappa
    c.use_pose, c.use_pose_resync, c.use_velocity, c.flow_aided = map(int, (cfg.use_pose, cfg.use_pose_resync, cfg.use_velocity, cfg.flow_aided))
    c.segm_delay, c.pose_delay = cfg.segm_delay, cfg.pose_delay
    return c


def _p(a):
    return None if a is None else a.ctypes.data


class CFilter:
    """Single-track ROFTFilter restatement (cpu_ref.cpp Filter)."""

    def __init__(self, cfg, p_mean0=None):
        self.cfg = cfg
        self.c = make_config(cfg)
        p0 = None if p_mean0 is None else np.ascontiguousarray(p_mean0, np.float64)
        self.h = lib().cref_filter_create(C.byref(self.c), _p(p0))
        self.flow_dtype = np.int16 if self.c.flow_s16 else np.float32

    def __del__(self):
        if getattr(self, "h", None):
            lib().cref_filter_destroy(self.h)
            self.h = None

    def step(self, depth, flow=None, mask=None, pose=None, dt=None):
        depth = np.ascontiguousarray(depth, np.float32)
        flow = None if flow is None else np.ascontiguousarray(flow, self.flow_dtype)
        mask = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        pose = None if pose is None else np.ascontiguousarray(pose, np.float64)
        lib().cref_filter_step(self.h, _p(depth), _p(flow), _p(mask), _p(pose), float(self.cfg.sample_time if dt is None else dt))

    def state(self):
        pm = np.empty(13); pc = np.empty((12, 12)); vm = np.empty(6); vc = np.empty((6, 6)); n = C.c_int32(0)
        lib().cref_filter_state(self.h, _p(pm), _p(pc), _p(vm), _p(vc), C.addressof(n))
        return pm, pc, vm, vc, n.value

    def mask(self):
        raw = np.zeros((self.cfg.height, self.cfg.width), np.uint8); thr = np.zeros_like(raw)
        lib().cref_filter_mask(self.h, _p(raw), _p(thr))
        return raw, thr


def flow_measurement(cfg, mask, depth, flow, dt):
    c = make_config(cfg)
    mask = np.ascontiguousarray(mask, np.uint8); depth = np.ascontiguousarray(depth, np.float32)
    flow = np.ascontiguousarray(flow, np.int16 if c.flow_s16 else np.float32)
    cap = int((mask != 0).sum()) + 1
    z = np.empty(2 * cap); Hm = np.empty((2 * cap, 6))
    n = lib().cref_flow_measurement(C.byref(c), _p(mask), _p(depth), _p(flow), float(dt), cap, _p(z), _p(Hm))
    return z[:2 * n].copy(), Hm[:2 * n].copy()


def skf_correct(cfg, x, P, z, Hm):
    c = make_config(cfg)
    x = np.ascontiguousarray(x, np.float64).copy(); P = np.ascontiguousarray(P, np.float64).copy()
    z = np.ascontiguousarray(z, np.float64); Hm = np.ascontiguousarray(Hm, np.float64)
    lib().cref_skf_correct(C.byref(c), _p(x), _p(P), _p(z), _p(Hm), z.shape[0] // 2)
    return x, P


def mask_warp(cfg, mask, flows, zero_origin):
    c = make_config(cfg)
    m = np.ascontiguousarray(mask, np.uint8).copy()
    fl = [np.ascontiguousarray(f, np.int16 if c.flow_s16 else np.float32) for f in flows]
    arr = (C.c_void_p * max(1, len(fl)))(*[f.ctypes.data for f in fl])
    lib().cref_mask_warp(C.byref(c), _p(m), arr, len(fl), int(zero_origin))
    return m


def ukf_predict(cfg, mean, cov, T):
    c = make_config(cfg)
    mean = np.ascontiguousarray(mean, np.float64).copy(); cov = np.ascontiguousarray(cov, np.float64).copy()
    lib().cref_ukf_predict(C.byref(c), _p(mean), _p(cov), float(T))
    return mean, cov


def ukf_correct(cfg, mean, cov, meas13, mtype):
    c = make_config(cfg)
    mean = np.ascontiguousarray(mean, np.float64).copy(); cov = np.ascontiguousarray(cov, np.float64).copy()
    meas13 = np.ascontiguousarray(meas13, np.float64)
    lib().cref_ukf_correct(C.byref(c), _p(mean), _p(cov), _p(meas13), int(mtype))
    return mean, cov


def timed_baseline(seq, stride: int, delay: int, seconds: float = 15.0, threads: Optional[int] = None, kind: str = "port",
                   max_tracks: int = 32, max_frames: int = 7):
    """Time the restated reference on a bounded sample of a synthetic sequence (roft_b200.synthetic).

    The sample is the first min(max_tracks, T) tracks x first min(max_frames, F) frames, replayed by
    `threads` workers (default: all host cores, one track at a time per worker) for ~`seconds`.
    """
    import roft_oracle as o
    T = min(int(seq.depth.shape[1]), max_tracks)
    F = min(int(seq.depth.shape[0]), max_frames)
    H, W = int(seq.depth.shape[2]), int(seq.depth.shape[3])
    ncpu = os.cpu_count() or 1
    threads = int(threads or ncpu)
    cfg = o.RoftConfig(width=W, height=H, subsampling_radius=float(stride), segm_delay=delay, pose_delay=delay,
                       flow_grid=seq.flow_grid, flow_scale=seq.flow_scale, sample_time=seq.dt)
    c = make_config(cfg)
    depth = np.ascontiguousarray(seq.depth[:F, :T].cpu().numpy())
    flow = np.ascontiguousarray(seq.flow[:F, :T].cpu().numpy())
    mask = np.ascontiguousarray(seq.mask[:F, :T].cpu().numpy())
    pose = np.ascontiguousarray(seq.pose[:F, :T].numpy())
    pv = np.ascontiguousarray(seq.pose_valid[:F, :T].numpy().astype(np.uint8))
    el = C.c_double(0.0)
    frames = lib().cref_timed_run(C.byref(c), T, F, threads, float(seconds), _p(depth), _p(flow), _p(mask), _p(pose), _p(pv),
                                  C.addressof(el), None)
    fps = frames / el.value if el.value > 0 else 0.0
    return {"value": fps, "unit": "tracked frames/s", "cores": threads, "host_cores": ncpu, "kind": kind,
            "ms_per_frame_per_core": 1e3 * threads / fps if fps else None,
            "sample": f"{T} tracks x {F} frames of the same synthetic workload (stride {stride}, delay {delay}), replayed for "
                      f"{el.value:.1f} s by {threads} threads; sequential per-pixel SKF as SKFCorrection.cpp:129-149; "
                      "lower bound on the real reference (no Eigen dynamic allocation, bfl::any copies or GL render)",
            "frames": int(frames), "seconds": el.value}
