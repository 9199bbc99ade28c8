"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on identical inputs.

Bars (BASELINE.json): warped / thresholded masks bit-exact; information matrix, information vector,
velocity and pose within 1e-4 relative; selected-pixel counts identical.
"""
import numpy as np
import pytest

import roft_oracle as o
from helpers import frame_inputs, quat_close, rel, sequence, small_cfg, to_roftb_config

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def api():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from roft_b200 import api as _api
    _api.load_library()
    return _api


def make_tracker(api, cfg, n=1, fmt="f32"):
    return api.Tracker(to_roftb_config(cfg, n, fmt))


# ------------------------------------------------------------------------------------------------
# mask synchronisation: bit-exact
# ------------------------------------------------------------------------------------------------
def _oracle_warp(mask, flows, cfg, zero_origin):
    m = mask.copy()
    if zero_origin:
        m[0, 0] = 0
    out = o.remap_exact(m, o.mask_warp_map(m, flows, cfg))
    return out, o.threshold_mask(out)


@pytest.mark.parametrize("fmt", ["f32", "s16"])
@pytest.mark.parametrize("mixed", [False, True])
def test_mask_sync_bit_exact(api, fmt, mixed):
    cfg = small_cfg(flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0, segm_delay=6)
    seq = sequence(cfg, 3, 8, flow_format=fmt, mixed_mask_values=mixed)
    trk = make_tracker(api, cfg, 3, fmt)
    # (a) single-flow propagation with zeroed origin, (b) new mask through 6 buffered flows
    for zero_origin, flows_idx in ((True, [3]), (False, [1, 2, 3, 4, 5, 6]), (False, [2])):
        masks = seq.mask[flows_idx[0] - 1].numpy()
        flows = [seq.flow[i].numpy() for i in flows_idx]
        raw, thr = trk.mask_sync(masks, flows, zero_origin)
        for t in range(3):
            eraw, ethr = _oracle_warp(masks[t], [f[t] for f in flows], cfg, zero_origin)
            assert np.array_equal(raw[t], eraw), (fmt, mixed, zero_origin, t, int((raw[t] != eraw).sum()))
            assert np.array_equal(thr[t], ethr)


def test_mask_sync_edge_cases(api):
    cfg = small_cfg(segm_delay=0)
    H, W = cfg.height, cfg.width
    rng = np.random.default_rng(0)
    trk = make_tracker(api, cfg, 1)
    cases = []
    # Q2: non-zero origin floods unmapped pixels in the new-mask branch
    m = np.zeros((H, W), np.uint8); m[40:90, 100:180] = 255; m[0, 0] = 255
    f = np.full((H, W, 2), 1.5, np.float32)
    cases.append((m, [f], False))
    cases.append((m, [f], True))
    # mixed values {1,2,255}, collisions (contracting flow), out-of-range / NaN / inf / huge flows
    m2 = rng.choice(np.array([0, 1, 2, 255], np.uint8), size=(H, W), p=[0.5, 0.1, 0.2, 0.2])
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    f2 = np.stack([(W / 2 - xx) * 0.3, (H / 2 - yy) * 0.3], -1).astype(np.float32)
    f2 += rng.normal(0, 0.7, f2.shape).astype(np.float32)
    bad = rng.random((H, W))
    f2[bad < 0.02] = np.nan
    f2[(bad >= 0.02) & (bad < 0.04)] = np.inf
    f2[(bad >= 0.04) & (bad < 0.06)] = 1e10
    f2[(bad >= 0.06) & (bad < 0.08)] = -3e9
    f3 = rng.normal(0, 30.0, f2.shape).astype(np.float32)  # large displacements leave the frame
    f4 = rng.uniform(-0.99, 0.99, f2.shape).astype(np.float32)  # t in (-1, 0) truncates to 0
    cases.append((m2, [f2], False))
    cases.append((m2, [f2, f3, f4], False))
    cases.append((m2, [f4], True))
    cases.append((m2, [], False))  # empty chain: identity map, unmapped -> src(0,0)
    cases.append((np.zeros((H, W), np.uint8), [f2], False))  # empty mask
    full = np.full((H, W), 7, np.uint8)
    cases.append((full, [f4, f4], False))
    for m, flows, zo in cases:
        raw, thr = trk.mask_sync(m[None], [f[None] for f in flows], zo)
        eraw, ethr = _oracle_warp(m, flows, cfg, zo)
        assert np.array_equal(raw[0], eraw), int((raw[0] != eraw).sum())
        assert np.array_equal(thr[0], ethr)


# ------------------------------------------------------------------------------------------------
# flow -> velocity measurement and Kalman correction
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt,stride,weight", [("f32", 1, True), ("f32", 35, True), ("f32", 7, False), ("s16", 1, True), ("s16", 5, True)])
def test_flow_velocity_normal_equations(api, fmt, stride, weight):
    cfg = small_cfg(flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0,
                    subsampling_radius=float(stride), weight_flow=weight)
    seq = sequence(cfg, 3, 3, flow_format=fmt)
    trk = make_tracker(api, cfg, 3, fmt)
    masks = o.threshold_mask(seq.mask[1].numpy().reshape(-1, cfg.width)).reshape(3, cfg.height, cfg.width)
    depth = seq.depth[1].numpy(); flow = seq.flow[2].numpy()
    xp = np.array([[0.02, -0.3, 0.05, -0.5, 0.2, -0.3]] * 3) * np.array([[1.0], [0.0], [-2.0]])
    dt = np.array([cfg.sample_time, 0.05, 0.02])
    lam, eta, cnt = trk.flow_velocity(masks, depth, flow, xp, dt)
    R = np.diag(cfg.cov_flow)
    for t in range(3):
        z, H, _ = o.flow_velocity_measurement(masks[t], depth[t], flow[t], cfg, dt[t])
        assert cnt[t] == z.shape[0] // 2
        _, _, Lm, em = o.skf_correct_information(xp[t], np.eye(6), z, H, R, weight)
        assert rel(lam[t], Lm) < TOL, (t, rel(lam[t], Lm))
        assert rel(eta[t], em) < TOL, (t, rel(eta[t], em))


def test_flow_velocity_full_mask_ragged_plane(api):
    """Every pixel selected on a 200 x 90 plane (140.625 units): the partial last unit carries candidates."""
    cfg = small_cfg(W=200, H=90, subsampling_radius=1.0, weight_flow=True)
    seq = sequence(cfg, 2, 3)
    trk = make_tracker(api, cfg, 2)
    masks = np.full((2, cfg.height, cfg.width), 255, np.uint8)
    depth = seq.depth[1].numpy(); flow = seq.flow[2].numpy()
    xp = np.array([[0.02, -0.3, 0.05, -0.5, 0.2, -0.3], [0.0] * 6])
    dt = np.array([cfg.sample_time, 0.05])
    lam, eta, cnt = trk.flow_velocity(masks, depth, flow, xp, dt)
    R = np.diag(cfg.cov_flow)
    for t in range(2):
        z, H, _ = o.flow_velocity_measurement(masks[t], depth[t], flow[t], cfg, dt[t])
        assert cnt[t] == z.shape[0] // 2
        _, _, Lm, em = o.skf_correct_information(xp[t], np.eye(6), z, H, R, True)
        assert rel(lam[t], Lm) < TOL, (t, rel(lam[t], Lm))
        assert rel(eta[t], em) < TOL, (t, rel(eta[t], em))


@pytest.mark.parametrize("stride", [1, 35])
def test_velocity_kf_matches_sequential_reference(api, stride):
    """Information-form GPU result vs the SEQUENTIAL per-pixel Kalman loop of SKFCorrection.cpp:129-149."""
    cfg = small_cfg(subsampling_radius=float(stride))
    seq = sequence(cfg, 2, 3)
    trk = make_tracker(api, cfg, 2)
    masks = seq.mask[1].numpy(); depth = seq.depth[1].numpy(); flow = seq.flow[2].numpy()
    x0 = np.array([[0.0] * 6, [0.05, -0.2, 0.0, -0.8, 0.1, 0.3]])
    P0 = np.stack([np.diag(cfg.v_cov0), np.diag(cfg.v_cov0) * 3.0])
    x, P, cnt = trk.velocity_kf(masks, depth, flow, x0, P0)
    Q = np.diag(list(cfg.v_sigma_linear) + list(cfg.v_sigma_angular))
    for t in range(2):
        z, H, _ = o.flow_velocity_measurement(masks[t], depth[t], flow[t], cfg, cfg.sample_time)
        xp, Pp = o.kf_predict(x0[t], P0[t], Q)
        xe, Pe = o.skf_correct(xp, Pp, z, H, np.diag(cfg.cov_flow), cfg.weight_flow)
        assert cnt[t] == z.shape[0] // 2
        assert rel(x[t], xe) < TOL and rel(P[t], Pe) < TOL, (rel(x[t], xe), rel(P[t], Pe))


def test_velocity_observability_gate_and_empty(api):
    cfg = small_cfg(subsampling_radius=1.0)
    H, W = cfg.height, cfg.width
    trk = make_tracker(api, cfg, 3)
    mask = np.zeros((3, H, W), np.uint8)
    mask[1, 10, 10:12] = 255         # 2 valid pixels < 3: unobservable
    mask[2, 20, 20:40] = 255         # enough pixels but all gated out by depth
    depth = np.full((3, H, W), 0.7, np.float32); depth[2] = 3.0
    flow = np.zeros((3, H, W, 2), np.float32)
    x0 = np.tile(np.arange(6.0), (3, 1)); P0 = np.tile(np.eye(6) * 0.01, (3, 1, 1))
    x, P, cnt = trk.velocity_kf(mask, depth, flow, x0, P0)
    assert list(cnt) == [0, 2, 0]
    assert np.array_equal(x, x0) and np.array_equal(P, P0)  # belief untouched (ROFTFilter.cpp:294-301)


def test_measurement_export_exact(api):
    cfg = small_cfg(subsampling_radius=3.0)
    seq = sequence(cfg, 1, 3)
    trk = make_tracker(api, cfg, 1)
    m = seq.mask[1, 0].numpy(); d = seq.depth[1, 0].numpy(); f = seq.flow[2, 0].numpy()
    z, H, n = trk.flow_measurement_export(m, d, f, cfg.sample_time)
    ze, He, _ = o.flow_velocity_measurement(m, d, f, cfg, cfg.sample_time)
    assert n == ze.shape[0] // 2
    assert np.array_equal(z, ze)  # FP32 quotients widened: exact
    assert np.allclose(H, He, rtol=1e-14, atol=0)


def test_masked_points_and_depth_l1(api):
    cfg = small_cfg()
    seq = sequence(cfg, 2, 2)
    trk = make_tracker(api, cfg, 2)
    m = seq.mask[1].numpy(); d = seq.depth[1].numpy()
    pts, cnt = trk.masked_points(m, d, max_depth=10.0)
    for t in range(2):
        e = o.masked_points(m[t], d[t], cfg, 10.0)
        assert cnt[t] == e.shape[0]
        assert np.allclose(pts[t, :cnt[t]], e, rtol=1e-14, atol=0)
    div = 4
    rng = np.random.default_rng(1)
    rend = np.where(rng.random((2, cfg.height // div, cfg.width // div)) < 0.7, 0.6 + 0.1 * rng.random((2, cfg.height // div, cfg.width // div)), 0.0).astype(np.float32)
    err, n = trk.masked_depth_l1(m, d, rend, div)
    for t in range(2):
        ee, en = o.masked_depth_l1(m[t], d[t], rend[t], div)
        assert n[t] == en
        assert abs(err[t] - ee) <= 1e-9 * max(1.0, abs(ee))


# ------------------------------------------------------------------------------------------------
# pose UKF
# ------------------------------------------------------------------------------------------------
def _rand_belief(rng, n):
    mean = np.zeros((n, 13)); cov = np.zeros((n, 12, 12))
    for i in range(n):
        mean[i, :9] = rng.normal(0, 0.3, 9); mean[i, 6:9] += [0, 0, 0.7]
        q = rng.normal(size=4); mean[i, 9:] = q / np.linalg.norm(q)
        B = rng.normal(size=(12, 12)) * 0.02
        cov[i] = B @ B.T + np.diag(rng.uniform(1e-4, 2e-3, 12))
    return mean, cov


def test_ukf_predict_and_correct(api):
    cfg = small_cfg()
    rng = np.random.default_rng(5)
    n = 6
    trk = make_tracker(api, cfg, 1)
    mean, cov = _rand_belief(rng, n)
    mean[0] = 0; mean[0, 9] = 1; cov[0] = np.diag(cfg.p_cov0)   # the exactly-degenerate initial belief
    dt = rng.uniform(0.02, 0.05, n)
    pm, pc = trk.ukf_predict(mean, cov, dt)
    for i in range(n):
        em, ec = o.ukf_predict(mean[i], cov[i], cfg, dt[i])
        assert rel(pm[i, :9], em[:9]) < TOL and quat_close(pm[i, 9:], em[9:]) < TOL
        assert rel(pc[i], ec) < TOL, rel(pc[i], ec)
    for mtype in (o.MEAS_VELOCITY, o.MEAS_POSE, o.MEAS_POSE_VELOCITY):
        meas = np.zeros((n, 13))
        meas[:, :9] = mean[:, :9] + rng.normal(0, 0.05, (n, 9))
        meas[:, :3] = mean[:, :3] + np.cross(mean[:, 3:6], -mean[:, 6:9]) + rng.normal(0, 0.05, (n, 3))
        dq = o.rotation_vector_to_quaternion(rng.normal(0, 0.05, (n, 3)))
        meas[:, 9:] = o.quat_mul(dq, mean[:, 9:])
        meas[1, 9:] *= -1  # measured quaternion on the other hemisphere
        cm, cc = trk.ukf_correct(pm, pc, meas, np.full(n, mtype, np.int32))
        for i in range(n):
            mv = {o.MEAS_VELOCITY: meas[i, :6], o.MEAS_POSE: meas[i, 6:], o.MEAS_POSE_VELOCITY: meas[i]}[mtype]
            em, ec = o.ukf_correct(pm[i], pc[i], mv, mtype, cfg)
            assert rel(cm[i, :9], em[:9]) < TOL, (mtype, i, rel(cm[i, :9], em[:9]))
            assert quat_close(cm[i, 9:], em[9:]) < TOL
            assert rel(cc[i], ec) < TOL, (mtype, i, rel(cc[i], ec))


# ------------------------------------------------------------------------------------------------
# the whole filter loop
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt,stride,resync", [("f32", 5, True), ("s16", 3, True), ("f32", 1, False)])
def test_filter_loop_matches_oracle(api, fmt, stride, resync):
    cfg = small_cfg(flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0,
                    subsampling_radius=float(stride), use_pose_resync=resync, segm_delay=4, pose_delay=4)
    _run_filter_loop(api, cfg, fmt, 3, 22)


def test_filter_loop_ragged_last_unit(api):
    """200 x 90 = 140.625 units of 128 px: the last unit of every plane is partial (bulk copies shorter than a stage,
    norm slots past the plane) on the TMA-staged path (dense flow, stride 1, weighting on), with delayed masks / poses."""
    cfg = small_cfg(W=200, H=90, subsampling_radius=1.0, use_pose_resync=True, segm_delay=3, pose_delay=3)
    _run_filter_loop(api, cfg, "f32", 2, 14)


def _run_filter_loop(api, cfg, fmt, T, F):
    seq = sequence(cfg, T, F, flow_format=fmt)
    x0 = np.zeros((T, 13)); x0[:, 6:] = seq.pose[0].numpy()
    trk = make_tracker(api, cfg, T, fmt)
    trk.init(x0)
    oracles = [o.RoftFilterOracle(cfg, x0[t]) for t in range(T)]
    for k in range(F):
        frs = [frame_inputs(seq, cfg, k, t) for t in range(T)]
        has_mask = frs[0].mask is not None
        mask = np.stack([f.mask for f in frs]) if has_mask else None
        pose = np.stack([f.pose if f.pose is not None else np.zeros(7) for f in frs])
        pv = np.array([f.pose is not None for f in frs], np.uint8)
        flow = np.stack([f.flow for f in frs]) if k > 0 else None
        trk.step(np.stack([f.depth for f in frs]), flow, mask, pose=pose, pose_valid=pv)
        pm, vm = trk.state()
        raw, thr = trk.mask()
        cnt, lam, eta = trk.velocity_info()
        for t in range(T):
            ep, ev = oracles[t].step(frs[t])
            orc = oracles[t]
            if orc.seg_source.mask is not None:
                assert np.array_equal(raw[t], orc.seg_source.mask), (k, t)
                assert np.array_equal(thr[t], orc.seg), (k, t)
            assert cnt[t] == orc.last_n_valid, (k, t, cnt[t], orc.last_n_valid)
            assert rel(vm[t], ev) < TOL or np.linalg.norm(vm[t] - ev) < 1e-9, (k, t, rel(vm[t], ev))
            assert rel(pm[t, :9], ep[:9]) < TOL, (k, t, rel(pm[t, :9], ep[:9]))
            assert quat_close(pm[t, 9:], ep[9:]) < TOL, (k, t)


def test_full_resolution_properties(api):
    """1280x720 (BASELINE size): oracle parity on one frame + size-independent properties."""
    cfg = o.RoftConfig(subsampling_radius=1.0)
    seq = sequence(cfg, 2, 3, target_coverage=0.4)
    trk = make_tracker(api, cfg, 2)
    m = seq.mask[1].numpy(); d = seq.depth[1].numpy(); f = seq.flow[2].numpy()
    xp = np.tile(np.array([0.02, -0.3, 0.05, -0.5, 0.2, -0.3]), (2, 1))
    lam, eta, cnt = trk.flow_velocity(m, d, f, xp)
    for t in range(2):
        z, H, _ = o.flow_velocity_measurement(m[t], d[t], f[t], cfg, cfg.sample_time)
        _, _, Lm, em = o.skf_correct_information(xp[t], np.eye(6), z, H, np.diag(cfg.cov_flow), True)
        assert cnt[t] == z.shape[0] // 2
        assert rel(lam[t], Lm) < TOL and rel(eta[t], em) < TOL
        assert np.allclose(lam[t], lam[t].T)  # symmetric
        assert np.all(np.linalg.eigvalsh(lam[t]) > 0)  # positive definite
    raw, thr = trk.mask_sync(m, [f], True)
    for t in range(2):
        eraw, ethr = _oracle_warp(m[t], [f[t]], cfg, True)
        assert np.array_equal(raw[t], eraw) and np.array_equal(thr[t], ethr)
    # zero flow: warp is the identity on the mask (idempotence), thresholding is idempotent
    raw0, thr0 = trk.mask_sync(m, [np.zeros_like(f)], True)
    m0 = m.copy(); m0[:, 0, 0] = 0
    assert np.array_equal(raw0, m0)
    assert np.array_equal(o.threshold_mask(thr0.reshape(-1, cfg.width)).reshape(thr0.shape), thr0)


@pytest.mark.parametrize("accum", [2, 1])
def test_full_resolution_filter_loop_vs_sequential_cpp(api, accum):
    """BASELINE size (1280x720, every masked pixel, ~3e5 measurements per frame): the whole filter loop against the
    C++ restatement with the SEQUENTIAL per-pixel Kalman update, for the default (auto) and the FP64 accumulation."""
    import cpu_ref
    cfg = o.RoftConfig(subsampling_radius=1.0, segm_delay=2, pose_delay=2)
    T, F = 2, 6
    seq = sequence(cfg, T, F, target_coverage=0.3)
    x0 = np.zeros((T, 13)); x0[:, 6:] = seq.pose[0].numpy()
    rc = to_roftb_config(cfg, T)
    rc.accum_fp64 = accum
    trk = api.Tracker(rc)
    trk.init(x0)
    refs = [cpu_ref.CFilter(cfg, x0[t]) for t in range(T)]
    for k in range(F):
        frs = [frame_inputs(seq, cfg, k, t) for t in range(T)]
        mask = np.stack([f.mask for f in frs]) if frs[0].mask is not None else None
        pose = np.stack([f.pose if f.pose is not None else np.zeros(7) for f in frs])
        pv = np.array([f.pose is not None for f in frs], np.uint8)
        flow = np.stack([f.flow for f in frs]) if k > 0 else None
        trk.step(np.stack([f.depth for f in frs]), flow, mask, pose=pose, pose_valid=pv)
        pm, vm = trk.state()
        raw, thr = trk.mask()
        cnt, _, _ = trk.velocity_info()
        for t in range(T):
            refs[t].step(frs[t].depth, frs[t].flow, frs[t].mask, frs[t].pose, frs[t].dt)
            epm, _, evm, _, en = refs[t].state()
            eraw, ethr = refs[t].mask()
            assert np.array_equal(raw[t], eraw) and np.array_equal(thr[t], ethr), (k, t)
            assert cnt[t] == en, (k, t, cnt[t], en)
            assert rel(vm[t], evm) < TOL or np.linalg.norm(vm[t] - evm) < 1e-9, (accum, k, t, rel(vm[t], evm))
            assert rel(pm[t, :9], epm[:9]) < TOL and quat_close(pm[t, 9:], epm[9:]) < TOL, (accum, k, t)


@pytest.mark.parametrize("fmt,stride,coverage,delay", [("f32", 1, 0.45, 4), ("s16", 35, 0.25, 6), ("f32", 35, 0.25, 6)])
def test_full_resolution_baseline_configs(api, fmt, stride, coverage, delay):
    """1280x720 at the parameters of BASELINE configs[4] (large masks >= 40 % of the frame, 4-frame mask-sync delay) and
    at the reference's own defaults (test/test.sh:63 nvof_1_slow = CV_16SC2 on a 4-px grid; config_fast_ycb.cfg:83
    subsampling_radius 35): the whole filter loop against the sequential C++ restatement, through two mask deliveries."""
    import cpu_ref
    cfg = o.RoftConfig(subsampling_radius=float(stride), segm_delay=delay, pose_delay=delay,
                       flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0)
    T, F = 2, 2 * delay + 2
    seq = sequence(cfg, T, F, target_coverage=coverage, flow_format=fmt)
    assert float((seq.mask[0] > 0).float().mean()) > 0.9 * coverage - 0.05
    x0 = np.zeros((T, 13)); x0[:, 6:] = seq.pose[0].numpy()
    trk = api.Tracker(to_roftb_config(cfg, T, fmt))
    trk.init(x0)
    refs = [cpu_ref.CFilter(cfg, x0[t]) for t in range(T)]
    for k in range(F):
        frs = [frame_inputs(seq, cfg, k, t) for t in range(T)]
        mask = np.stack([f.mask for f in frs]) if frs[0].mask is not None else None
        pose = np.stack([f.pose if f.pose is not None else np.zeros(7) for f in frs])
        pv = np.array([f.pose is not None for f in frs], np.uint8)
        flow = np.stack([f.flow for f in frs]) if k > 0 else None
        trk.step(np.stack([f.depth for f in frs]), flow, mask, pose=pose, pose_valid=pv)
        pm, vm = trk.state()
        raw, thr = trk.mask()
        cnt, _, _ = trk.velocity_info()
        for t in range(T):
            refs[t].step(frs[t].depth, frs[t].flow, frs[t].mask, frs[t].pose, frs[t].dt)
            epm, _, evm, _, en = refs[t].state()
            eraw, ethr = refs[t].mask()
            assert np.array_equal(raw[t], eraw) and np.array_equal(thr[t], ethr), (k, t)
            assert cnt[t] == en, (k, t, cnt[t], en)
            assert rel(vm[t], evm) < TOL or np.linalg.norm(vm[t] - evm) < 1e-9, (fmt, stride, k, t, rel(vm[t], evm))
            assert rel(pm[t, :9], epm[:9]) < TOL and quat_close(pm[t, 9:], epm[9:]) < TOL, (fmt, stride, k, t)


def test_batch_of_many_tracks_matches_single_track_runs(api):
    """The batch is only a batch: 24 tracks stepped together (several clusters in flight, scratch slots reused, tracks
    scheduled largest first) give bit-identical beliefs and masks to the same tracks stepped one context at a time."""
    cfg = small_cfg(subsampling_radius=1.0, segm_delay=3, pose_delay=3)
    T, F = 24, 9
    seq = sequence(cfg, T, F)
    x0 = np.zeros((T, 13)); x0[:, 6:] = seq.pose[0].numpy()

    def run(tracks):
        n = len(tracks)
        trk = make_tracker(api, cfg, n)
        trk.init(x0[tracks])
        outs = []
        for k in range(F):
            frs = [frame_inputs(seq, cfg, k, t) for t in tracks]
            mask = np.stack([f.mask for f in frs]) if frs[0].mask is not None else None
            pose = np.stack([f.pose if f.pose is not None else np.zeros(7) for f in frs])
            pv = np.array([f.pose is not None for f in frs], np.uint8)
            trk.step(np.stack([f.depth for f in frs]), np.stack([f.flow for f in frs]) if k > 0 else None, mask, pose=pose,
                     pose_valid=pv)
            pm, vm = trk.state()
            raw, _ = trk.mask(thresholded=False)
            outs.append((pm.copy(), vm.copy(), raw.copy()))
        return outs

    batch = run(list(range(T)))
    for t in (0, 7, 23):
        single = run([t])
        for k in range(F):
            assert np.array_equal(batch[k][0][t], single[k][0][0]), (t, k)
            assert np.array_equal(batch[k][1][t], single[k][1][0]), (t, k)
            assert np.array_equal(batch[k][2][t], single[k][2][0]), (t, k)


def test_ho3d_format_single_track(api):
    """BASELINE configs[2]: HO-3D-format 640x480 single track (SURVEY 8d intrinsics fx = fy = 617, cx = 312, cy = 241),
    every masked pixel, masks and poses delayed by 4 frames - parity of the whole filter loop against the sequential
    C++ restatement, and the per-frame latency of the C-ABI step with HOST buffers (upload + step + read-back)."""
    import time
    import cpu_ref
    cfg = o.RoftConfig(width=640, height=480, fx=617.0, fy=617.0, cx=312.0, cy=241.0, subsampling_radius=1.0,
                       segm_delay=4, pose_delay=4)
    T, F = 1, 14
    seq = sequence(cfg, T, F, target_coverage=0.2)
    x0 = np.zeros((T, 13)); x0[:, 6:] = seq.pose[0].numpy()
    trk = api.Tracker(to_roftb_config(cfg, T))
    trk.init(x0)
    ref = cpu_ref.CFilter(cfg, x0[0])
    lat = []
    for k in range(F):
        fr = frame_inputs(seq, cfg, k, 0)
        mask = fr.mask[None] if fr.mask is not None else None
        pose = (fr.pose if fr.pose is not None else np.zeros(7))[None]
        pv = np.array([fr.pose is not None], np.uint8)
        flow = fr.flow[None] if k > 0 else None
        t0 = time.perf_counter()
        trk.step(fr.depth[None], flow, mask, pose=pose, pose_valid=pv)
        pm, vm = trk.state()
        lat.append((time.perf_counter() - t0) * 1e3)
        raw, thr = trk.mask()
        cnt, _, _ = trk.velocity_info()
        ref.step(fr.depth, fr.flow, fr.mask, fr.pose, fr.dt)
        epm, _, evm, _, en = ref.state()
        eraw, ethr = ref.mask()
        assert np.array_equal(raw[0], eraw) and np.array_equal(thr[0], ethr), k
        assert cnt[0] == en, (k, cnt[0], en)
        assert rel(vm[0], evm) < TOL or np.linalg.norm(vm[0] - evm) < 1e-9, (k, rel(vm[0], evm))
        assert rel(pm[0, :9], epm[:9]) < TOL and quat_close(pm[0, 9:], epm[9:]) < TOL, k
    lat = np.sort(np.array(lat[2:]))
    print(f"\nHO-3D-format single track, host buffers: p50 {lat[len(lat) // 2]:.3f} ms, max {lat[-1]:.3f} ms per frame")
    assert np.isfinite(lat).all()


def test_worklist_diagnostics(api):
    """roftb_get_worklist: non-empty 128-px units and segmentation pixels (byte > 1) of the mask the last step used."""
    cfg = small_cfg(subsampling_radius=1.0, segm_delay=2, pose_delay=2)
    T = 2
    seq = sequence(cfg, T, 4)
    x0 = np.zeros((T, 13)); x0[:, 6:] = seq.pose[0].numpy()
    trk = make_tracker(api, cfg, T)
    trk.init(x0)
    orc = [o.RoftFilterOracle(cfg, x0[t]) for t in range(T)]
    for k in range(4):
        frs = [frame_inputs(seq, cfg, k, t) for t in range(T)]
        prev_seg = [None if orc[t].seg is None else orc[t].seg.copy() for t in range(T)]
        mask = np.stack([f.mask for f in frs]) if frs[0].mask is not None else None
        pose = np.stack([f.pose if f.pose is not None else np.zeros(7) for f in frs])
        pv = np.array([f.pose is not None for f in frs], np.uint8)
        trk.step(np.stack([f.depth for f in frs]), np.stack([f.flow for f in frs]) if k > 0 else None, mask, pose=pose,
                 pose_valid=pv)
        units, pixels = trk.worklist()
        for t in range(T):
            orc[t].step(frs[t])
            if prev_seg[t] is None:
                continue
            # the velocity pass of step k reads the segmentation synchronised at step k-1
            flat = (prev_seg[t].reshape(-1) > 1)
            assert pixels[t] == int(flat.sum()), (k, t)
            pad = (-flat.size) % 128
            u = np.pad(flat, (0, pad)).reshape(-1, 128).any(axis=1).sum()
            assert units[t] == int(u), (k, t)


# ---- f1: render-and-compare pose outlier rejection -------------------------------------------------------------------
def _render_close(got, exp):
    """Depth tiles agree: same silhouette up to a handful of edge pixels (the model matrix goes through the C library's
    cosf / sinf on one side and numpy's on the other: positions can differ by one 1/256-px step), depth within 1e-4."""
    both = (got > 0) & (exp > 0)
    mism = int(((got > 0) != (exp > 0)).sum())
    assert mism <= max(2, 0.002 * int((exp > 0).sum())), (mism, int((exp > 0).sum()))
    assert both.sum() > 0 and np.abs(got[both] - exp[both]).max() <= 1e-4 * exp[both].max()


@pytest.mark.gpu
def test_depth_rasteriser_matches_oracle(lib_built):
    """roftb_render_depth (SICAD::superimpose depth, SICAD.cpp:924-1066) against the numpy restatement: cuboid and a
    2048-triangle sphere, poses incl. partially out of frame and behind the camera, dividers 1 / 2 / 4."""
    from roft_b200 import api
    from roft_b200.synthetic import cuboid_mesh
    from test_oracle import _octasphere
    cfg = small_cfg()
    trk = api.Tracker(to_roftb_config(cfg, 1))
    rng = np.random.default_rng(5)
    for verts, faces in (cuboid_mesh([0.08, 0.105, 0.03]), _octasphere(0.07, 4)):
        trk.set_mesh(verts, faces)
        poses = []
        for k in range(6):
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
            poses.append(np.concatenate([[rng.uniform(-0.05, 0.05), rng.uniform(-0.03, 0.03), rng.uniform(0.4, 0.9)], ax, [rng.uniform(0, 3.1)]]))
        poses.append(np.array([0.12, 0.05, 0.5, 0, 0, 1, 0.3]))      # partly outside the frame
        poses.append(np.array([0.0, 0.0, -0.5, 1, 0, 0, 0.0]))       # behind the camera: nothing
        poses.append(np.array([0.0, 0.0, 0.6, 0, 0, 0, 0.0]))        # zero axis (identity rotation)
        poses = np.array(poses)
        for div in (1, 2, 4):
            got = trk.render_depth(poses, div)
            assert got.shape == (len(poses), cfg.height // div, cfg.width // div)
            for i, p in enumerate(poses):
                exp = o.render_depth(verts, faces, p, cfg.width // div, cfg.height // div, cfg.fx / div, cfg.fy / div, cfg.cx / div, cfg.cy / div)
                if not exp.any():
                    assert not got[i].any()
                else:
                    _render_close(got[i], exp)


@pytest.mark.gpu
def test_pick_best_alternative_matches_oracle(lib_built):
    """roftb_pick_best_alternative (ROFTFilter.cpp:467-621): same selection and likelihoods as the oracle for a
    consistent / displaced pair in both orders, a near-tie, and an empty mask (DBL_MAX, first alternative kept)."""
    from roft_b200 import api
    from roft_b200.synthetic import cuboid_mesh
    cfg = small_cfg()
    T = 3
    seq = sequence(cfg, T, 3, corrupt=False)
    verts, faces = cuboid_mesh(seq.half[0].numpy())
    trk = api.Tracker(to_roftb_config(cfg, T))
    trk.set_mesh(verts, faces)
    k = 2
    masks = seq.mask[k].numpy().copy(); depths = seq.depth[k].numpy()
    alts = np.zeros((T, 2, 13))
    for t in range(T):
        alts[t, :, 6:] = seq.gt_pose[k, t].numpy()
    alts[0, 1, 8] += 0.12          # track 0: second alternative displaced -> first wins
    alts[1, 0, 6] += 0.05          # track 1: first displaced -> second wins
    alts[2, 1, 8] += 0.002         # track 2: near tie -> first stays
    for div, gain in ((2, 0.01), (4, 1.0)):
        sel, lik = trk.pick_best_alternative(masks, depths, alts, div, gain)
        for t in range(T):
            es, el = o.pick_best_alternative(cfg, verts, faces, [alts[t, 0], alts[t, 1]], masks[t], depths[t], div, gain)
            assert sel[t] == es, (t, sel[t], es, lik[t], el)
            assert np.allclose(lik[t], el, rtol=2e-3), (t, lik[t], el)
        assert list(sel) == [0, 1, 0]
    masks[0] = 0
    sel, lik = trk.pick_best_alternative(masks, depths, alts, 2, 0.01)
    assert sel[0] == 0 and lik[0, 0] == np.finfo(np.float64).max and lik[0, 1] == np.finfo(np.float64).max


@pytest.mark.gpu
@pytest.mark.parametrize("resync", [True, False])
def test_filter_loop_with_outlier_rejection_matches_oracle(api, resync):
    """cfg.outlier_rejection: every 13-sized pose measurement goes through correct_outlier_rejection (two UKF corrections,
    depth render of both, masked L1 on the buffered features, choice - ROFTFilter.cpp:346-359, 649-676) on the device.
    Some delivered poses are displaced by 6-10 cm so that both outcomes occur; beliefs match the oracle frame by frame
    and the oracle took each branch at least once."""
    import torch
    from roft_b200.synthetic import cuboid_mesh
    cfg = small_cfg(subsampling_radius=2.0, use_pose_resync=resync, segm_delay=3, pose_delay=3,
                    outlier_rejection=True, outlier_rejection_divider=2)
    T, F = 2, 26
    seq = sequence(cfg, T, F)
    # gross pose outliers on some deliveries (the delayed source delivers frame k - delay at steps k % delay == 0)
    for f_idx, t, off in ((6, 0, (0.08, 0.0, 0.0)), (12, 1, (0.0, -0.06, 0.05)), (15, 0, (0.0, 0.0, 0.10))):
        seq.pose[f_idx, t, :3] += torch.tensor(off, dtype=seq.pose.dtype)
    verts, faces = cuboid_mesh(seq.half[0].numpy())
    x0 = np.zeros((T, 13)); x0[:, 6:] = seq.pose[0].numpy()
    trk = make_tracker(api, cfg, T, "f32")
    if resync:
        trk.set_mesh(verts, faces)
    else:  # batched extension: one unit mesh, per-track scale (roftb_set_mesh_scale)
        trk.set_mesh(*cuboid_mesh([1.0, 1.0, 1.0]))
        trk.set_mesh_scale(np.stack([seq.half[t].numpy() for t in range(T)]))
    trk.init(x0)
    oracles = [o.RoftFilterOracle(cfg, x0[t], mesh=cuboid_mesh(seq.half[t].numpy())) for t in range(T)]
    for k in range(F):
        frs = [frame_inputs(seq, cfg, k, t) for t in range(T)]
        has_mask = frs[0].mask is not None
        mask = np.stack([f.mask for f in frs]) if has_mask else None
        pose = np.stack([f.pose if f.pose is not None else np.zeros(7) for f in frs])
        pv = np.array([f.pose is not None for f in frs], np.uint8)
        flow = np.stack([f.flow for f in frs]) if k > 0 else None
        trk.step(np.stack([f.depth for f in frs]), flow, mask, pose=pose, pose_valid=pv)
        pm, vm = trk.state()
        for t in range(T):
            ep, ev = oracles[t].step(frs[t])
            assert rel(vm[t], ev) < TOL or np.linalg.norm(vm[t] - ev) < 1e-9, (k, t, rel(vm[t], ev))
            assert rel(pm[t, :9], ep[:9]) < TOL, (k, t, rel(pm[t, :9], ep[:9]), oracles[t].or_selected)
            assert quat_close(pm[t, 9:], ep[9:]) < TOL, (k, t)
    picks = sum((orc.or_selected for orc in oracles), [])
    assert 0 in picks and 1 in picks, picks
    # the same run without a host synchronisation between the steps (the pose stream trails, the buffered features go
    # through their staging copies while later steps recycle the planes): same final beliefs
    trk2 = make_tracker(api, cfg, T, "f32")
    trk2.set_mesh(verts, faces)
    if not resync:
        trk2.set_mesh(*cuboid_mesh([1.0, 1.0, 1.0]))
        trk2.set_mesh_scale(np.stack([seq.half[t].numpy() for t in range(T)]))
    trk2.init(x0)
    for k in range(F):
        frs = [frame_inputs(seq, cfg, k, t) for t in range(T)]
        mask = np.stack([f.mask for f in frs]) if frs[0].mask is not None else None
        pose = np.stack([f.pose if f.pose is not None else np.zeros(7) for f in frs])
        pv = np.array([f.pose is not None for f in frs], np.uint8)
        flow = np.stack([f.flow for f in frs]) if k > 0 else None
        trk2.step(np.stack([f.depth for f in frs]), flow, mask, pose=pose, pose_valid=pv)
    pm2, vm2 = trk2.state()
    assert rel(pm2, pm) < 1e-9 and rel(vm2, vm) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("parts", [2, 3])
def test_pipelined_part_contexts_match_oracle(api, monkeypatch, parts):
    """ROFTB_PARTS: a context run as 2-3 complete part contexts over consecutive pieces of the track range (the default from
    64 tracks: one part's launch tail overlaps the other's bulk).  Same frame-by-frame parity as the single context, with
    and without the render-and-compare test, and every per-track read-back (state, masks, counts) in track order."""
    import torch
    from roft_b200.synthetic import cuboid_mesh
    monkeypatch.setenv("ROFTB_PARTS", str(parts))
    cfg = small_cfg(subsampling_radius=2.0, use_pose_resync=True, segm_delay=3, pose_delay=3)
    _run_filter_loop(api, cfg, "f32", 5, 12)
    cfg = small_cfg(subsampling_radius=2.0, use_pose_resync=True, segm_delay=3, pose_delay=3, outlier_rejection=True,
                    outlier_rejection_divider=2)
    T, F = 4, 14
    seq = sequence(cfg, T, F)
    seq.pose[6, 3, :3] += torch.tensor([0.08, 0.0, 0.0], dtype=seq.pose.dtype)
    verts, faces = cuboid_mesh(seq.half[0].numpy())
    x0 = np.zeros((T, 13)); x0[:, 6:] = seq.pose[0].numpy()
    trk = make_tracker(api, cfg, T, "f32")
    trk.set_mesh(verts, faces)
    trk.init(x0)
    oracles = [o.RoftFilterOracle(cfg, x0[t], mesh=(verts, faces)) for t in range(T)]
    for k in range(F):
        frs = [frame_inputs(seq, cfg, k, t) for t in range(T)]
        mask = np.stack([f.mask for f in frs]) if frs[0].mask is not None else None
        pose = np.stack([f.pose if f.pose is not None else np.zeros(7) for f in frs])
        pv = np.array([f.pose is not None for f in frs], np.uint8)
        flow = np.stack([f.flow for f in frs]) if k > 0 else None
        trk.step(np.stack([f.depth for f in frs]), flow, mask, pose=pose, pose_valid=pv)
        pm, vm = trk.state()
        units, pixels = trk.worklist()
        assert units.shape == (T,) and (k < 2 or (units > 0).all())
        for t in range(T):
            ep, ev = oracles[t].step(frs[t])
            assert rel(vm[t], ev) < TOL or np.linalg.norm(vm[t] - ev) < 1e-9, (k, t)
            assert rel(pm[t, :9], ep[:9]) < TOL and quat_close(pm[t, 9:], ep[9:]) < TOL, (k, t, oracles[t].or_selected)
    assert 1 in sum((orc.or_selected for orc in oracles), [])
