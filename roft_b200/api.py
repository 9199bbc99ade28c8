"""ctypes binding of libroft_b200.so (include/roft_b200.h).

The Python layer is only a thin caller of the C ABI - the product is the CUDA library.  There is
NO CPU fallback: loading fails loudly when the shared library is missing, and creating a
context fails loudly when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libroft_b200.so")
_lib = None

FLOW_F32 = 13
FLOW_S16 = 11
MEM_HOST = 0
MEM_DEVICE = 1
MEAS_NONE, MEAS_VELOCITY, MEAS_POSE, MEAS_POSE_VELOCITY = 0, 1, 2, 3
MAX_DELAY = 8

EXPORTED_SYMBOLS = [
    "roftb_config_default", "roftb_create", "roftb_destroy", "roftb_last_error", "roftb_sync", "roftb_join", "roftb_version",
    "roftb_kernel_launches", "roftb_stream", "roftb_profile", "roftb_filter_init", "roftb_filter_step", "roftb_get_state",
    "roftb_get_mask", "roftb_get_velocity_info", "roftb_get_worklist", "roftb_mask_sync", "roftb_flow_velocity", "roftb_velocity_kf",
    "roftb_flow_measurement_export", "roftb_masked_points", "roftb_masked_depth_l1", "roftb_ukf_predict",
    "roftb_ukf_correct", "roftb_set_mesh", "roftb_set_mesh_scale", "roftb_render_depth", "roftb_pick_best_alternative",
]


class RoftbConfig(C.Structure):
    _fields_ = [
        ("n_tracks", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("sample_time", C.c_double),
        ("flow_format", C.c_int32), ("flow_grid", C.c_int32), ("flow_scale", C.c_float),
        ("cov_flow", C.c_double * 2), ("depth_maximum", C.c_double),
        ("subsampling_radius", C.c_int32), ("weight_flow", C.c_int32),
        ("v_sigma", C.c_double * 6), ("v_cov0", C.c_double * 6),
        ("p_sigma_linear", C.c_double * 3), ("p_sigma_angular", C.c_double * 3), ("p_cov0", C.c_double * 12),
        ("cov_v", C.c_double * 3), ("cov_w", C.c_double * 3), ("cov_x", C.c_double * 3), ("cov_q", C.c_double * 3),
        ("ut_alpha", C.c_double), ("ut_beta", C.c_double), ("ut_kappa", C.c_double),
        ("use_pose", C.c_int32), ("use_pose_resync", C.c_int32), ("use_velocity", C.c_int32), ("flow_aided", C.c_int32),
        ("segm_delay", C.c_int32), ("pose_delay", C.c_int32),
        ("device", C.c_int32), ("accum_fp64", C.c_int32),
        ("outlier_rejection", C.c_int32), ("outlier_rejection_divider", C.c_int32), ("outlier_rejection_gain", C.c_double),
    ]


class RoftbFrame(C.Structure):
    _fields_ = [
        ("memory", C.c_int32),
        ("depth", C.c_void_p), ("depth_track_stride", C.c_int64),
        ("flow", C.c_void_p), ("flow_track_stride", C.c_int64),
        ("mask", C.c_void_p), ("mask_track_stride", C.c_int64),
        ("flow_valid", C.c_void_p), ("mask_valid", C.c_void_p),
        ("pose", C.c_void_p), ("pose_valid", C.c_void_p), ("dt", C.c_void_p),
    ]


class RoftbError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """Load libroft_b200.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RoftbError(f"{_LIB_PATH} is missing: build it with `python -m roft_b200.build` "
                         "(roft_b200 has no CPU fallback)")
    lib = C.CDLL(_LIB_PATH)
    lib.roftb_last_error.restype = C.c_char_p
    lib.roftb_last_error.argtypes = [C.c_void_p]
    lib.roftb_create.argtypes = [C.POINTER(RoftbConfig), C.POINTER(C.c_void_p)]
    lib.roftb_destroy.argtypes = [C.c_void_p]
    lib.roftb_destroy.restype = None
    lib.roftb_config_default.argtypes = [C.POINTER(RoftbConfig)]
    lib.roftb_config_default.restype = None
    lib.roftb_sync.argtypes = [C.c_void_p]
    lib.roftb_join.argtypes = [C.c_void_p]
    lib.roftb_kernel_launches.argtypes = [C.c_void_p]
    lib.roftb_kernel_launches.restype = C.c_int64
    lib.roftb_stream.argtypes = [C.c_void_p]
    lib.roftb_stream.restype = C.c_void_p
    lib.roftb_profile.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.roftb_filter_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.roftb_filter_step.argtypes = [C.c_void_p, C.POINTER(RoftbFrame)]
    lib.roftb_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    lib.roftb_get_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.roftb_get_velocity_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.roftb_get_worklist.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.roftb_mask_sync.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.roftb_flow_velocity.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 8
    lib.roftb_velocity_kf.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 7
    lib.roftb_flow_measurement_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32,
                                                  C.c_void_p, C.c_void_p, C.c_void_p]
    lib.roftb_masked_points.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]
    lib.roftb_masked_depth_l1.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.roftb_set_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    lib.roftb_set_mesh_scale.argtypes = [C.c_void_p, C.c_void_p]
    lib.roftb_render_depth.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.roftb_pick_best_alternative.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double,
                                                C.c_void_p, C.c_void_p]
    lib.roftb_ukf_predict.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.roftb_ukf_correct.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def default_config(**overrides) -> RoftbConfig:
    """config/config_fast_ycb.cfg defaults, with keyword overrides (tuples for array fields)."""
    lib = load_library()
    cfg = RoftbConfig()
    lib.roftb_config_default(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(f"roftb_config has no field {k}")
        cur = getattr(cfg, k)
        if isinstance(cur, C.Array):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(cfg, k, v)
    return cfg


def _np(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _ptr(a) -> Optional[int]:
    """Raw address of a numpy array or a torch tensor (host or device)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


class Tracker:
    """Batched ROFT filter over n_tracks independent object tracks on one GPU.

    Mirrors ROFTFilter (src/roft-lib/src/ROFTFilter.cpp): ``init`` = initialization_step,
    ``step`` = filtering_step for every track, ``state`` reads the beliefs back.
    """

    def __init__(self, cfg: RoftbConfig):
        self._lib = load_library()
        self.cfg = cfg
        h = C.c_void_p()
        rc = self._lib.roftb_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise RoftbError(f"roftb_create failed ({rc}): {self._lib.roftb_last_error(None).decode()}")
        self._h = h
        self.n_tracks = cfg.n_tracks
        self.W, self.H = cfg.width, cfg.height
        self.Wf, self.Hf = cfg.width // cfg.flow_grid, cfg.height // cfg.flow_grid
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.roftb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc < 0:
            raise RoftbError(f"{what} failed ({rc}): {self._lib.roftb_last_error(self._h).decode()}")
        return rc

    # ---- filter loop -------------------------------------------------------------------
    def init(self, p_mean0=None, v_mean0=None):
        p = _np(p_mean0, np.float64) if p_mean0 is not None else None
        v = _np(v_mean0, np.float64) if v_mean0 is not None else None
        self._check(self._lib.roftb_filter_init(self._h, _ptr(p), _ptr(v)), "roftb_filter_init")

    def step(self, depth, flow=None, mask=None, *, flow_valid=None, mask_valid=None, pose=None, pose_valid=None,
             dt=None, device: bool = False):
        """One filtering_step for all tracks.

        depth [T,H,W] f32, flow [T,Hf,Wf,2] f32/i16 or None, mask [T,H,W] u8 or None; numpy arrays
        (host path, device=False) or torch CUDA tensors (zero-copy, device=True).  The small per-track
        arrays are numpy / lists.  Device tensors must outlive the next max(segm_delay,1) steps.
        """
        T = self.n_tracks
        fr = RoftbFrame()
        fr.memory = MEM_DEVICE if device else MEM_HOST
        keep = []
        if not device:
            depth = _np(depth, np.float32)
            flow = None if flow is None else _np(flow, np.int16 if self.cfg.flow_format == FLOW_S16 else np.float32)
            mask = None if mask is None else _np(mask, np.uint8)
        keep += [depth, flow, mask]
        fr.depth = _ptr(depth)
        fr.depth_track_stride = self.H * self.W
        fr.flow = _ptr(flow)
        fr.flow_track_stride = self.Hf * self.Wf * 2
        fr.mask = _ptr(mask)
        fr.mask_track_stride = self.H * self.W

        def small(a, dtype, n):
            if a is None:
                return None
            a = _np(a, dtype).reshape(-1)
            assert a.size == n, (a.size, n)
            keep.append(a)
            return a.ctypes.data

        fr.flow_valid = small(flow_valid, np.uint8, T)
        fr.mask_valid = small(mask_valid, np.uint8, T)
        fr.pose = small(pose, np.float64, T * 7)
        fr.pose_valid = small(pose_valid, np.uint8, T)
        fr.dt = small(dt, np.float64, T)
        self._check(self._lib.roftb_filter_step(self._h, C.byref(fr)), "roftb_filter_step")
        # device tensors must stay alive while later steps may still read them
        self._keep.append(keep)
        if len(self._keep) > 2 * MAX_DELAY + 4:
            self._keep.pop(0)

    def state(self, cov: bool = False):
        T = self.n_tracks
        pm = np.empty((T, 13)); vm = np.empty((T, 6))
        pc = np.empty((T, 12, 12)) if cov else None
        vc = np.empty((T, 6, 6)) if cov else None
        self._check(self._lib.roftb_get_state(self._h, _ptr(pm), _ptr(pc), _ptr(vm), _ptr(vc)), "roftb_get_state")
        return (pm, vm, pc, vc) if cov else (pm, vm)

    def mask(self, raw: bool = True, thresholded: bool = True):
        T = self.n_tracks
        r = np.empty((T, self.H, self.W), np.uint8) if raw else None
        t = np.empty((T, self.H, self.W), np.uint8) if thresholded else None
        self._check(self._lib.roftb_get_mask(self._h, _ptr(r), _ptr(t)), "roftb_get_mask")
        return r, t

    def velocity_info(self):
        T = self.n_tracks
        cnt = np.empty(T, np.int32); lam = np.empty((T, 6, 6)); eta = np.empty((T, 6))
        self._check(self._lib.roftb_get_velocity_info(self._h, _ptr(cnt), _ptr(lam), _ptr(eta)), "roftb_get_velocity_info")
        return cnt, lam, eta

    def worklist(self):
        """(non-empty 128-px units, segmentation pixels) per track of the last step's synchronised mask."""
        T = self.n_tracks
        units = np.empty(T, np.int32); pixels = np.empty(T, np.int32)
        self._check(self._lib.roftb_get_worklist(self._h, _ptr(units), _ptr(pixels)), "roftb_get_worklist")
        return units, pixels

    def sync(self):
        self._check(self._lib.roftb_sync(self._h), "roftb_sync")

    def join(self):
        """Make the main stream wait for the internal streams (call before recording an end-of-region event)."""
        self._check(self._lib.roftb_join(self._h), "roftb_join")

    PHASES = ("prep", "flow_pass_a", "median_select", "flow_pass_b", "epilogue", "mask_scatter", "ukf")

    def profile(self, enable: bool):
        """Read the per-phase device times (ms per step) gathered so far, then enable/disable profiling."""
        ms = np.zeros(7); n = C.c_int64(0)
        self._check(self._lib.roftb_profile(self._h, int(enable), _ptr(ms), C.addressof(n)), "roftb_profile")
        return dict(zip(self.PHASES, ms.tolist())), n.value

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.roftb_kernel_launches(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.roftb_stream(self._h) or 0)

    # ---- stateless operators -------------------------------------------------------------
    def _flow_np(self, flow):
        return _np(flow, np.int16 if self.cfg.flow_format == FLOW_S16 else np.float32)

    def mask_sync(self, mask, flows: Sequence, zero_origin: bool):
        """mask [N,H,W] u8, flows: list of n_flows arrays [N,Hf,Wf,2] (oldest first) -> (raw, thresholded)."""
        mask = _np(mask, np.uint8)
        N = mask.shape[0]
        fl = self._flow_np(np.stack(flows, 0)) if len(flows) else None
        raw = np.empty_like(mask); thr = np.empty_like(mask)
        self._check(self._lib.roftb_mask_sync(self._h, N, _ptr(mask), _ptr(fl), len(flows), int(zero_origin),
                                              _ptr(raw), _ptr(thr)), "roftb_mask_sync")
        return raw, thr

    def flow_velocity(self, mask, depth, flow, x_pred=None, dt=None):
        mask = _np(mask, np.uint8); depth = _np(depth, np.float32); flow = self._flow_np(flow)
        N = mask.shape[0]
        xp = _np(x_pred, np.float64) if x_pred is not None else np.zeros((N, 6))
        dtv = _np(dt, np.float64) if dt is not None else None
        lam = np.empty((N, 6, 6)); eta = np.empty((N, 6)); cnt = np.empty(N, np.int32)
        self._check(self._lib.roftb_flow_velocity(self._h, N, _ptr(mask), _ptr(depth), _ptr(flow), _ptr(xp), _ptr(dtv),
                                                  _ptr(lam), _ptr(eta), _ptr(cnt)), "roftb_flow_velocity")
        return lam, eta, cnt

    def velocity_kf(self, mask, depth, flow, x, P, dt=None):
        mask = _np(mask, np.uint8); depth = _np(depth, np.float32); flow = self._flow_np(flow)
        N = mask.shape[0]
        x = _np(x, np.float64).copy(); P = _np(P, np.float64).copy()
        dtv = _np(dt, np.float64) if dt is not None else None
        cnt = np.empty(N, np.int32)
        self._check(self._lib.roftb_velocity_kf(self._h, N, _ptr(mask), _ptr(depth), _ptr(flow), _ptr(dtv), _ptr(x), _ptr(P),
                                                _ptr(cnt)), "roftb_velocity_kf")
        return x, P, cnt

    def flow_measurement_export(self, mask, depth, flow, dt: float, capacity: Optional[int] = None):
        mask = _np(mask, np.uint8); depth = _np(depth, np.float32); flow = self._flow_np(flow)
        cap = int(capacity if capacity is not None else self.H * self.W)
        z = np.empty(cap * 2); Hm = np.empty((cap * 2, 6)); n = C.c_int32(0)
        self._check(self._lib.roftb_flow_measurement_export(self._h, _ptr(mask), _ptr(depth), _ptr(flow), float(dt), cap,
                                                            _ptr(z), _ptr(Hm), C.addressof(n)), "roftb_flow_measurement_export")
        k = min(n.value, cap)
        return z[:2 * k].copy(), Hm[:2 * k].copy(), n.value

    def masked_points(self, mask, depth, max_depth: float = 10.0, capacity: Optional[int] = None):
        mask = _np(mask, np.uint8); depth = _np(depth, np.float32)
        N = mask.shape[0]
        cap = int(capacity if capacity is not None else self.H * self.W)
        pts = np.empty((N, cap, 3)); cnt = np.empty(N, np.int32)
        self._check(self._lib.roftb_masked_points(self._h, N, _ptr(mask), _ptr(depth), float(max_depth), cap, _ptr(pts),
                                                  _ptr(cnt)), "roftb_masked_points")
        return pts, cnt

    def masked_depth_l1(self, mask, depth, rendered, divider: int):
        mask = _np(mask, np.uint8); depth = _np(depth, np.float32); rendered = _np(rendered, np.float32)
        N = mask.shape[0]
        err = np.empty(N); cnt = np.empty(N, np.int32)
        self._check(self._lib.roftb_masked_depth_l1(self._h, N, _ptr(mask), _ptr(depth), _ptr(rendered), int(divider),
                                                    _ptr(err), _ptr(cnt)), "roftb_masked_depth_l1")
        return err, cnt

    def set_mesh(self, vertices, faces):
        v = _np(vertices, np.float32); f = _np(faces, np.int32)
        self._check(self._lib.roftb_set_mesh(self._h, _ptr(v), int(v.shape[0]), _ptr(f), int(f.shape[0])), "roftb_set_mesh")

    def set_mesh_scale(self, scale):
        sc = None if scale is None else _np(scale, np.float32)
        self._check(self._lib.roftb_set_mesh_scale(self._h, _ptr(sc)), "roftb_set_mesh_scale")

    def render_depth(self, poses7, divider: int):
        p = _np(poses7, np.float64)
        N = p.shape[0]
        out = np.empty((N, self.cfg.height // divider, self.cfg.width // divider), np.float32)
        self._check(self._lib.roftb_render_depth(self._h, N, _ptr(p), int(divider), _ptr(out)), "roftb_render_depth")
        return out

    def pick_best_alternative(self, segmentation, depth, alternatives, divider: int, gain: float):
        m = _np(segmentation, np.uint8); d = _np(depth, np.float32); a = _np(alternatives, np.float64)
        N = m.shape[0]
        sel = np.empty(N, np.int32); lik = np.empty((N, 2))
        self._check(self._lib.roftb_pick_best_alternative(self._h, N, _ptr(m), _ptr(d), _ptr(a), int(divider), C.c_double(gain),
                                                          _ptr(sel), _ptr(lik)), "roftb_pick_best_alternative")
        return sel, lik

    def ukf_predict(self, mean, cov, dt=None):
        mean = _np(mean, np.float64).copy(); cov = _np(cov, np.float64).copy()
        N = mean.shape[0]
        dtv = _np(dt, np.float64) if dt is not None else None
        self._check(self._lib.roftb_ukf_predict(self._h, N, _ptr(mean), _ptr(cov), _ptr(dtv)), "roftb_ukf_predict")
        return mean, cov

    def ukf_correct(self, mean, cov, meas, meas_type):
        mean = _np(mean, np.float64).copy(); cov = _np(cov, np.float64).copy()
        meas = _np(meas, np.float64); mt = _np(meas_type, np.int32)
        N = mean.shape[0]
        self._check(self._lib.roftb_ukf_correct(self._h, N, _ptr(mean), _ptr(cov), _ptr(meas), _ptr(mt)), "roftb_ukf_correct")
        return mean, cov
