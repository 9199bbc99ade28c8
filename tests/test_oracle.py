"""CPU tests of the oracle itself (no GPU): the numpy/cv2 restatement against literal loop transcriptions,
the C++ restatement (oracle/cpu_ref.cpp) against the numpy one, and mathematical properties of the
UPSTREAM-RECALL pieces."""
import math

import numpy as np
import pytest

import cpu_ref
import roft_oracle as o
from helpers import frame_inputs, quat_close, rel, sequence, small_cfg

cv2 = pytest.importorskip("cv2")


# ---- OpenCV semantics the bit-exact mask path relies on ------------------------------
def test_cv2_semantics():
    rng = np.random.default_rng(0)
    m = rng.integers(0, 256, (37, 52), dtype=np.uint8)
    m[rng.random(m.shape) < 0.5] = 0
    pts = cv2.findNonZero(m).reshape(-1, 2)
    ys, xs = np.nonzero(m)  # row-major
    assert np.array_equal(pts[:, 0], xs) and np.array_equal(pts[:, 1], ys)
    # integer-coordinate INTER_LINEAR remap is an exact gather, in-place safe, unmapped -> src(0,0)
    mp = np.zeros(m.shape + (2,), np.float32)
    sel = rng.random(m.shape) < 0.6
    mp[sel, 0] = rng.integers(0, m.shape[1], sel.sum())
    mp[sel, 1] = rng.integers(0, m.shape[0], sel.sum())
    out = cv2.remap(m, mp, None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    exp = m[mp[..., 1].astype(int), mp[..., 0].astype(int)]
    assert np.array_equal(out, exp)
    assert np.all(out[~sel] == m[0, 0])
    t = cv2.threshold(np.arange(256, dtype=np.uint8).reshape(16, 16), 1, 255, cv2.THRESH_BINARY)[1].ravel()
    assert t[0] == 0 and t[1] == 0 and np.all(t[2:] == 255)


def _literal_map(mask, flows, grid, scale, D):
    """Line-by-line transcription of ImageSegmentationOFAidedSource<T>::map (hpp:235-281), pure Python loops."""
    H, W = mask.shape
    mp = np.zeros((H, W, 2), np.float32)
    start = 0
    if D > 0:
        start = max(0, len(flows) - D)
    f32 = np.float32

    def cint(t):
        t = f32(t)
        if not np.isfinite(t) or abs(float(t)) >= 2147483648.0:
            return -2147483648
        return int(float(t))

    ys, xs = np.nonzero(mask)
    for py, px in zip(ys, xs):
        tx, ty = f32(px), f32(py)
        error = False
        for j in range(start, len(flows)):
            if cint(tx) < 0 or cint(tx) >= W or cint(ty) < 0 or cint(ty) >= H:
                error = True
                break
            with np.errstate(all="ignore"):
                f = flows[j][cint(ty / f32(grid)), cint(tx / f32(grid))]
                tx = f32(tx + f32(f32(f[0]) / f32(scale)))
                ty = f32(ty + f32(f32(f[1]) / f32(scale)))
        if error or cint(tx) < 0 or cint(tx) >= W or cint(ty) < 0 or cint(ty) >= H:
            continue
        mp[cint(ty), cint(tx)] = (px, py)
    return mp


@pytest.mark.parametrize("fmt", ["f32", "s16"])
def test_mask_warp_matches_literal_transcription(fmt):
    rng = np.random.default_rng(1)
    H, W = 36, 48
    grid, scale = (1, 1.0) if fmt == "f32" else (4, 32.0)
    cfg = small_cfg(W, H, flow_grid=grid, flow_scale=scale, segm_delay=3)
    mask = rng.choice(np.array([0, 1, 2, 255], np.uint8), size=(H, W), p=[0.5, 0.1, 0.2, 0.2])
    flows = []
    for _ in range(4):
        if fmt == "f32":
            f = rng.normal(0, 2.5, (H, W, 2)).astype(np.float32)
            bad = rng.random((H, W))
            f[bad < 0.03] = np.nan
            f[(bad > 0.03) & (bad < 0.05)] = np.inf
            f[(bad > 0.05) & (bad < 0.07)] = -4e9
        else:
            f = rng.integers(-150, 150, (H // 4, W // 4, 2)).astype(np.int16)
        flows.append(f)
    lit = _literal_map(mask, flows, grid, scale, cfg.segm_delay)
    vec = o.mask_warp_map(mask, flows, cfg)
    assert np.array_equal(lit, vec)
    exp = o.remap_exact(mask, vec)
    for zo in (False, True):
        m = mask.copy()
        if zo:
            m[0, 0] = 0
        e = o.remap_exact(m, o.mask_warp_map(m, flows, cfg))
        assert np.array_equal(cpu_ref.mask_warp(cfg, mask, flows, zo), e)
    assert exp.shape == mask.shape


def test_delay_schedule_matches_cpp_modulo():
    # DatasetImageSegmentationDelayed.cpp:42-63 with C++ '%' semantics
    def literal(head, delay, head0=0):
        index = head - delay
        rem = int(math.fmod(index - head0, delay))  # C++ truncating remainder
        if rem != 0:
            return None
        return head0 if index < 0 else index
    s = o.DelayedMaskSchedule(6)
    assert [s.index_for(h) for h in range(20)] == [literal(h, 6) for h in range(20)]
    assert s.index_for(0) == 0 and s.index_for(6) == 0 and s.index_for(12) == 6 and s.index_for(7) is None
    assert o.DelayedMaskSchedule(0).index_for(5) == 5


# ---- velocity path -----------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt,stride", [("f32", 1), ("f32", 35), ("s16", 4)])
def test_flow_measurement_numpy_vs_cpp(fmt, stride):
    cfg = small_cfg(flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0, subsampling_radius=float(stride))
    seq = sequence(cfg, 1, 3, flow_format=fmt)
    m = o.threshold_mask(seq.mask[1, 0].numpy()); d = seq.depth[1, 0].numpy(); f = seq.flow[2, 0].numpy()
    z, H, uv = o.flow_velocity_measurement(m, d, f, cfg, 0.031)
    zc, Hc = cpu_ref.flow_measurement(cfg, m, d, f, 0.031)
    assert z.shape[0] > 20 and np.array_equal(z, zc)
    assert np.allclose(H, Hc, rtol=1e-15, atol=0)
    # gates: no selected pixel has invalid flow / depth
    dd = d[uv[:, 1], uv[:, 0]]
    assert np.all((dd > 0) & (dd < cfg.depth_maximum)) and np.all(np.isfinite(z)) and np.all(np.abs(z) < 1e9)


def test_sequential_skf_equals_information_form_and_cpp():
    cfg = small_cfg(subsampling_radius=9.0)
    seq = sequence(cfg, 1, 3)
    m = seq.mask[1, 0].numpy(); d = seq.depth[1, 0].numpy(); f = seq.flow[2, 0].numpy()
    z, H, _ = o.flow_velocity_measurement(m, d, f, cfg, cfg.sample_time)
    xp = np.array([0.01, -0.2, 0.03, -0.4, 0.1, 0.2]); Pp = np.eye(6) * 0.101
    R = np.diag(cfg.cov_flow)
    for weighting in (True, False):
        c2 = small_cfg(subsampling_radius=9.0, weight_flow=weighting)
        xs, Ps = o.skf_correct(xp, Pp, z, H, R, weighting)
        xi, Pi, _, _ = o.skf_correct_information(xp, Pp, z, H, R, weighting)
        xc, Pc = cpu_ref.skf_correct(c2, xp, Pp, z, H)
        assert rel(xi, xs) < 1e-10 and rel(Pi, Ps) < 1e-10
        assert rel(xc, xs) < 1e-12 and rel(Pc, Ps) < 1e-12


def _skf_weights_transcribed(innov):
    """SKFCorrection.cpp:91-116 line by line, with Eigen's semantics spelled out in numpy: MatrixXd and
    Map<MatrixXd> are COLUMN-major (``order="F"``), ``rowwise().norm()`` is the 2-norm of each row."""
    sub = 2                                                                     # measurement_sub_size_
    rows = innov.shape[0]
    innovation_vector = innov[:rows].reshape((rows // sub, sub), order="F")      # :93  Map<MatrixXd>(data, rows/2, 2)
    norms = np.sqrt((innovation_vector ** 2).sum(axis=1))                       # :94  rowwise().norm()
    norms = np.sort(norms)                                                      # :95  std::sort
    mi = norms[norms.size // 2]                                                 # :98
    if norms.size % 2 == 0:                                                     # :99-100
        mi = 0.5 * (norms[norms.size // 2 - 1] + norms[norms.size // 2])
    b = np.abs(norms - mi).sum() / norms.size                                   # :102
    lik = np.ones(norms.size)                                                   # :104
    if b > 1e-4:                                                                # :106
        for j in range(norms.size):                                             # :108-112
            seg = innov[j * sub:j * sub + sub]                                  #      col(0).segment(j*2, 2)
            lik[j] = max(1.0 / (2 * b) * math.exp(-abs(np.sqrt((seg ** 2).sum()) - mi) / b), 1e-6)
        lik /= lik.max()                                                        # :114
    return lik, mi, b


def test_laplacian_weights_reference_rules():
    # SKFCorrection.cpp:95-116: even/odd median, b <= 1e-4 -> all ones, floor 1e-6, max-normalised.
    # Q3: the median / b come from the COLUMN-major view of the interleaved innovations: r_i = |(nu[i], nu[N+i])|
    nu = np.array([0.3, -0.4, 0.1, 0.2, -0.6, 0.8, 0.05, 0.0])           # 4 pixels
    r = np.sqrt(nu[:4] ** 2 + nu[4:] ** 2)
    assert np.allclose(o.laplacian_stat_norms(nu), r)
    s = np.sort(r); m = 0.5 * (s[1] + s[2]); b = np.abs(r - m).mean()
    pix = np.sqrt(nu[0::2] ** 2 + nu[1::2] ** 2)
    e = np.maximum(np.exp(-np.abs(pix - m) / b) / (2 * b), 1e-6)
    assert np.allclose(o.laplacian_likelihoods(nu), e / e.max(), rtol=1e-15)
    # it is NOT the median of the per-pixel norms
    assert abs(m - np.median(pix)) > 1e-2
    # identical innovations: b = 0 -> all ones
    assert np.all(o.laplacian_likelihoods(np.tile([0.3, 0.4], 5)) == 1.0)
    far = o.laplacian_likelihoods(np.array([0.1, 0.0, 0.1, 0.0, 0.1, 0.0, 0.1001, 0.0, 0.1, 0.0, 50.0, 0.0, 0.1002, 0.0]))
    assert far.max() == 1.0 and far.min() > 0


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 64, 257, 1000])
def test_laplacian_weights_match_line_by_line_transcription(n):
    """Odd and even N (the column-major pairing differs: for odd N a dx is paired with a dy of another pixel)."""
    rng = np.random.default_rng(n)
    nu = rng.laplace(0.0, 0.4, 2 * n) + rng.normal(0, 0.05, 2 * n)
    nu[rng.integers(0, 2 * n, max(1, n // 10))] += 25.0  # outliers
    lik, mi, b = _skf_weights_transcribed(nu)
    assert np.allclose(o.laplacian_likelihoods(nu), lik, rtol=1e-14, atol=0)
    # explicit element pairing for both parities: r_i = |(nu[i], nu[N + i])|
    r = np.array([math.hypot(nu[i], nu[n + i]) for i in range(n)])
    assert np.allclose(np.sort(o.laplacian_stat_norms(nu)), np.sort(r), rtol=1e-15)
    # and the C++ restatement agrees through the full correction
    if n >= 3:
        H = rng.normal(size=(2 * n, 6)); xp = rng.normal(size=6) * 0.1; Pp = np.eye(6) * 0.1
        z = nu + H @ xp
        cfg = small_cfg(weight_flow=True)
        xs, Ps = o.skf_correct(xp, Pp, z, H, np.diag(cfg.cov_flow), True)
        xc, Pc = cpu_ref.skf_correct(cfg, xp, Pp, z, H)
        assert rel(xc, xs) < 1e-10 and rel(Pc, Ps) < 1e-10


# ---- bfl pieces (UPSTREAM-RECALL): properties --------------------------------------------------------
def test_ut_weights_and_linear_map_exactness():
    for n in (18, 21, 24):
        wm, wc, c = o.ut_weights(n, 1.0, 2.0, 0.0)
        assert abs(wm.sum() - 1) < 1e-15 and wm[0] == 0 and wc[0] == 2 and c == n
    rng = np.random.default_rng(2)
    B = rng.normal(size=(12, 12)) * 0.05
    P = B @ B.T + np.eye(12) * 1e-3
    A = o.cov_sqrt(P)
    assert np.allclose(A @ A.T, P, atol=1e-15)
    As = o.cov_sqrt(P, "svd")  # same sigma set up to column order / sign => same outer product
    assert np.allclose(As @ As.T, P, atol=1e-15)
    w, V = o.jacobi_eigh(np.eye(6) * 1e-3)  # degenerate: untouched
    assert np.array_equal(V, np.eye(6))


def test_quaternion_boxplus_boxminus():
    rng = np.random.default_rng(3)
    q = rng.normal(size=(20, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    r = rng.normal(0, 0.4, (20, 3))
    q2 = o.sum_quaternion_rotation_vector(q[0], r)  # exp(r) (x) q
    back = o.diff_quaternion(q2, q[0])
    assert np.allclose(back, r, atol=1e-12)
    assert np.allclose(o.diff_quaternion(-q2, q[0]), r, atol=1e-12)  # short way round: sign of q is irrelevant
    # left-multiplication convention = CartesianQuaternionModel.cpp:111-122
    w = np.array([0.3, -0.2, 0.5]); T = 0.04
    sk = np.zeros((4, 4)); sk[0, 1:] = -w; sk[1:, 0] = w
    sk[1, 2], sk[1, 3], sk[2, 1], sk[2, 3], sk[3, 1], sk[3, 2] = -w[2], w[1], w[2], -w[0], -w[1], w[0]
    nw = np.linalg.norm(w)
    ref = (math.cos(nw * T / 2) * np.eye(4) + math.sin(nw * T / 2) / nw * sk) @ q[1]
    assert np.allclose(o.sum_quaternion_rotation_vector(q[1], w * T)[0], ref, atol=1e-15)


def _belief(rng):
    mean = np.zeros(13); mean[:9] = rng.normal(0, 0.3, 9); mean[8] += 0.7
    q = rng.normal(size=4); mean[9:] = q / np.linalg.norm(q)
    B = rng.normal(size=(12, 12)) * 0.02
    return mean, B @ B.T + np.diag(rng.uniform(1e-4, 2e-3, 12))


def test_ukf_numpy_vs_cpp_and_zero_rate_limit():
    cfg = small_cfg()
    rng = np.random.default_rng(4)
    for _ in range(3):
        mean, cov = _belief(rng)
        pm, pc = o.ukf_predict(mean, cov, cfg, 0.033)
        cm, cc = cpu_ref.ukf_predict(cfg, mean, cov, 0.033)
        assert rel(cm[:9], pm[:9]) < 1e-12 and quat_close(cm[9:], pm[9:]) < 1e-12 and rel(cc, pc) < 1e-10
        # linear part of the motion model is exact under the UT: x' = x + v T, v' = v
        assert np.allclose(pm[6:9], mean[6:9] + mean[0:3] * 0.033, atol=1e-12) and np.allclose(pm[:6], mean[:6], atol=1e-12)
        for mtype in (o.MEAS_VELOCITY, o.MEAS_POSE, o.MEAS_POSE_VELOCITY):
            meas = np.zeros(13)
            meas[:6] = pm[:6] + rng.normal(0, 0.05, 6)
            meas[6:9] = pm[6:9] + rng.normal(0, 0.01, 3)
            meas[9:] = o.sum_quaternion_rotation_vector(pm[9:], rng.normal(0, 0.05, 3))[0]
            mv = {o.MEAS_VELOCITY: meas[:6], o.MEAS_POSE: meas[6:], o.MEAS_POSE_VELOCITY: meas}[mtype]
            em, ec = o.ukf_correct(pm, pc, mv, mtype, cfg)
            fm, fc = cpu_ref.ukf_correct(cfg, pm, pc, meas, mtype)
            assert rel(fm[:9], em[:9]) < 1e-10 and quat_close(fm[9:], em[9:]) < 1e-10 and rel(fc, ec) < 1e-9
            assert np.all(np.linalg.eigvalsh(0.5 * (ec + ec.T)) > 0)
            assert np.trace(ec) < np.trace(pc)  # a measurement reduces uncertainty


@pytest.mark.parametrize("fmt", ["f32", "s16"])
def test_filter_loop_numpy_vs_cpp(fmt):
    cfg = small_cfg(flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0,
                    subsampling_radius=3.0, segm_delay=3, pose_delay=3)
    seq = sequence(cfg, 1, 12, flow_format=fmt)
    x0 = np.zeros(13); x0[6:] = seq.pose[0, 0].numpy()
    a = o.RoftFilterOracle(cfg, x0); b = cpu_ref.CFilter(cfg, x0)
    for k in range(12):
        fr = frame_inputs(seq, cfg, k, 0)
        ep, ev = a.step(fr)
        b.step(fr.depth, fr.flow, fr.mask, fr.pose, fr.dt)
        pm, pc, vm, vc, n = b.state()
        raw, thr = b.mask()
        assert n == a.last_n_valid
        assert np.array_equal(raw, a.seg_source.mask) and np.array_equal(thr, a.seg)
        assert rel(vm, ev) < 1e-6 or np.linalg.norm(vm - ev) < 1e-12
        assert rel(pm[:9], ep[:9]) < 1e-6 and quat_close(pm[9:], ep[9:]) < 1e-6
    # the filter tracks: velocity estimate close to the ground-truth twist at the camera origin
    gt = seq.gt_twist[0].numpy()
    v_o = gt[:3] + np.cross(gt[3:], -seq.gt_pose[11, 0, :3].numpy())
    assert np.linalg.norm(ev[3:] - gt[3:]) < 0.25 * max(1.0, np.linalg.norm(gt[3:]))
    assert np.linalg.norm(ev[:3] - v_o) < 0.25 * max(0.3, np.linalg.norm(v_o))


def test_masked_points_and_l1_small():
    cfg = small_cfg(64, 48)
    rng = np.random.default_rng(5)
    mask = (rng.random((48, 64)) < 0.4).astype(np.uint8) * 255
    depth = rng.uniform(-0.1, 2.5, (48, 64)).astype(np.float32)
    pts = o.masked_points(mask, depth, cfg, 2.0)
    ys, xs = np.nonzero(mask)
    ok = (depth[ys, xs] > 0) & (depth[ys, xs] < 2.0)
    assert pts.shape[0] == ok.sum()
    assert np.allclose(pts[:, 2], depth[ys, xs][ok])
    rend = np.where(rng.random((24, 32)) < 0.5, 0.8, 0.0).astype(np.float32)
    err, n = o.masked_depth_l1(mask, depth, rend, 2)
    e2, n2 = 0.0, 0
    for k, (y, x) in enumerate(zip(ys, xs)):
        if k % 2:
            continue
        dd, r = depth[y, x], rend[y // 2, x // 2]
        if dd > 0 and dd < 2.0 and r != 0:
            e2 += float(abs(np.float32(dd) - np.float32(r))); n2 += 1
    assert n == n2 and abs(err - e2) < 1e-9


# ---- f1: render-and-compare pose outlier rejection -------------------------------------------------------------------
def _octasphere(radius, level=3):
    """Subdivided octahedron: a closed mesh with many small triangles."""
    v = [np.array(p, np.float64) for p in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1))]
    f = [(0, 2, 4), (2, 1, 4), (1, 3, 4), (3, 0, 4), (2, 0, 5), (1, 2, 5), (3, 1, 5), (0, 3, 5)]
    for _ in range(level):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return (np.array(v) * radius).astype(np.float32), np.array(f, np.int32)


def test_oracle_rasteriser_against_closed_forms():
    """SICAD depth restatement: a sphere renders the analytic ray/sphere depth at the pixel centres (u + 0.5, v + 0.5), the
    silhouette is the analytic disc up to the faceting, and the quaternion -> axis-angle conversion is Eigen's."""
    W, H, fx, fy, cx, cy = 160, 120, 200.0, 210.0, 80.0, 60.0
    r, c = 0.08, np.array([0.03, -0.02, 0.6])
    verts, faces = _octasphere(r, 4)
    q = np.array([0.9, 0.1, -0.3, 0.2]); q /= np.linalg.norm(q)
    aa = o.quaternion_to_axis_angle(q)
    assert abs(np.linalg.norm(aa[:3]) - 1) < 1e-12 and abs(np.cos(aa[3] / 2) - q[0]) < 1e-12
    assert np.allclose(o.quaternion_to_axis_angle(-q), aa, atol=1e-12)   # q and -q: same axis, same angle
    assert np.array_equal(o.quaternion_to_axis_angle(np.array([1.0, 0, 0, 0])), [1.0, 0, 0, 0])
    d = o.render_depth(verts, faces, np.concatenate([c, aa]), W, H, fx, fy, cx, cy)
    uu, vv = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    ray = np.stack([(uu - cx) / fx, (vv - cy) / fy, np.ones_like(uu)], -1)
    a = (ray * ray).sum(-1); b = -2 * (ray @ c); cc = c @ c - r * r
    disc = b * b - 4 * a * cc
    z = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), 0.0)
    hit = d > 0
    assert abs(int(hit.sum()) - int((disc > 0).sum())) < 0.03 * (disc > 0).sum()     # faceted silhouette
    inner = hit & (disc > 0.5 * disc.max())
    assert inner.sum() > 500 and np.abs(d[inner] - z[inner]).max() < 5e-4               # chord error of the facets
    # a fronto-parallel square: exact coverage and constant depth
    sq = np.array([[-0.1, -0.05, 0], [0.1, -0.05, 0], [0.1, 0.05, 0], [-0.1, 0.05, 0]], np.float32)
    d2 = o.render_depth(sq, np.array([[0, 1, 2], [0, 2, 3]], np.int32), np.array([0, 0, 0.5, 1, 0, 0, 0.0]), W, H, fx, fy, cx, cy)
    u0, u1 = cx - 0.1 * fx / 0.5, cx + 0.1 * fx / 0.5
    v0, v1 = cy - 0.05 * fy / 0.5, cy + 0.05 * fy / 0.5
    exp = ((uu >= u0) & (uu <= u1) & (vv >= v0) & (vv <= v1))
    assert np.array_equal(d2 > 0, exp) and np.abs(d2[exp] - 0.5).max() < 5e-5
    # a triangle behind the camera is dropped, not wrapped around
    assert not o.render_depth(sq, np.array([[0, 1, 2]], np.int32), np.array([0, 0, -0.5, 1, 0, 0, 0.0]), W, H, fx, fy, cx, cy).any()


def test_oracle_pick_best_alternative_prefers_the_consistent_pose():
    from roft_b200.synthetic import cuboid_mesh
    cfg = small_cfg()
    seq = sequence(cfg, 1, 2, corrupt=False)
    verts, faces = cuboid_mesh(seq.half[0].numpy())
    good = np.zeros(13); good[6:] = seq.gt_pose[1, 0].numpy()
    bad = good.copy(); bad[8] += 0.15
    sel, lik = o.pick_best_alternative(cfg, verts, faces, [good, bad], seq.mask[1, 0].numpy(), seq.depth[1, 0].numpy(), 2, 0.01)
    assert sel == 0 and lik[0] < 0.5 * lik[1]
    sel, lik = o.pick_best_alternative(cfg, verts, faces, [bad, good], seq.mask[1, 0].numpy(), seq.depth[1, 0].numpy(), 2, 0.01)
    assert sel == 1
    # no samples (empty mask): DBL_MAX for both, the first alternative stays (ROFTFilter.cpp:569-583)
    sel, lik = o.pick_best_alternative(cfg, verts, faces, [good, bad], np.zeros_like(seq.mask[1, 0].numpy()), seq.depth[1, 0].numpy(), 2, 0.01)
    assert sel == 0 and lik[0] == np.finfo(np.float64).max
