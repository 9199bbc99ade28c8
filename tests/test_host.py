"""Host-side C++ mirror of the reference interface (roft_b200/host): dataset formats, delay schedule and - on a GPU -
the batched ROFTFilter driven from Fast-YCB-format directories by the roft_b200_tracker executable."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import roft_oracle as o
from helpers import frame_inputs, quat_close, rel, sequence, small_cfg
from roft_b200 import dataset_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "roft_b200", "host")


@pytest.fixture(scope="module")
def hostlib(lib_built):
    subprocess.check_call(["make", "-s", "-C", HOST])
    return ctypes.CDLL(os.path.join(HOST, "libroft_b200_host.so"))


def test_flow_file_format_roundtrip(hostlib, tmp_path):
    rng = np.random.default_rng(0)
    for flow in (rng.normal(0, 3, (18, 32, 2)).astype(np.float32), rng.integers(-900, 900, (9, 16, 2)).astype(np.int16)):
        a, b = str(tmp_path / "a.float"), str(tmp_path / "b.float")
        dataset_io.write_flow(a, flow)
        t = ctypes.c_int(0); c = ctypes.c_ulonglong(0); r = ctypes.c_ulonglong(0)
        assert hostlib.rofth_flow_roundtrip(a.encode(), b.encode(), ctypes.byref(t), ctypes.byref(c), ctypes.byref(r)) == 0
        assert (t.value, c.value, r.value) == (11 if flow.dtype == np.int16 else 13, flow.shape[1], flow.shape[0])
        assert open(a, "rb").read() == open(b, "rb").read()  # 20-byte header, no padding (OpticalFlowUtilities.cpp:38-62)
        assert os.path.getsize(a) == 20 + flow.nbytes
        assert np.array_equal(dataset_io.read_flow(b), flow)
    assert hostlib.rofth_flow_roundtrip(b"/nonexistent.float", b"/tmp/x", ctypes.byref(t), ctypes.byref(c), ctypes.byref(r)) < 0


def test_delay_schedule_and_flow_validity(hostlib):
    for delay in (1, 4, 6):
        s = o.DelayedMaskSchedule(delay)
        for head in range(0, 30):
            e = s.index_for(head)
            assert hostlib.rofth_delivered_index(head, delay, 0, 1) == (-1 if e is None else e)
    hostlib.rofth_is_flow_valid.argtypes = [ctypes.c_float, ctypes.c_float]
    for fx, fy, ok in ((0.5, -3.0, 1), (float("nan"), 0.0, 0), (0.0, float("inf"), 0), (1e10, 0.0, 0), (-9.9e8, 9.9e8, 1)):
        assert hostlib.rofth_is_flow_valid(fx, fy) == ok


def test_depth_file_roundtrip(tmp_path):
    d = np.random.default_rng(1).uniform(0, 2, (12, 20)).astype(np.float32)
    p = str(tmp_path / "0.float")
    dataset_io.write_depth(p, d)
    assert os.path.getsize(p) == 16 + d.nbytes and np.array_equal(dataset_io.read_depth(p), d)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["f32", "s16"])
def test_tracker_executable_matches_oracle(hostlib, tmp_path, fmt):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = small_cfg(flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0, subsampling_radius=4.0,
                    segm_delay=3, pose_delay=3, sample_time=1.0 / 30.0)
    T, F = 2, 10
    seq = sequence(cfg, T, F, flow_format=fmt, target_coverage=0.3)
    args = [os.path.join(HOST, "roft_b200_tracker"), "--log", str(tmp_path), "--stride", "4", "--desired-fps", "10"]
    for t in range(T):
        root = str(tmp_path / f"seq{t}")
        dataset_io.write_sequence(root, seq, t, fx=cfg.fx, fy=cfg.fy, cx=cfg.cx, cy=cfg.cy)
        args += ["--sequence", root]
    out = subprocess.run(args, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"tracked {F} frames x {T} tracks" in out.stdout
    for t in range(T):
        pose_log = np.loadtxt(tmp_path / f"track{t}_pose_estimate.txt")
        vel_log = np.loadtxt(tmp_path / f"track{t}_velocity_estimate.txt")
        assert pose_log.shape == (F, 13) and vel_log.shape == (F, 6)
        x0 = np.zeros(13)
        # the executable initialises from the first pose row, which went through axis-angle text
        aa = dataset_io.quat_to_axis_angle(seq.pose[0, t].numpy()[3:])
        x0[6:9] = seq.pose[0, t, :3].numpy()
        x0[9] = np.cos(aa[3] / 2); x0[10:] = np.sin(aa[3] / 2) * aa[:3]
        orc = o.RoftFilterOracle(cfg, x0)
        for k in range(F):
            fr = frame_inputs(seq, cfg, k, t)
            if fr.pose is not None:  # poses.txt stores axis-angle
                a2 = dataset_io.quat_to_axis_angle(fr.pose[3:])
                fr.pose = np.concatenate([fr.pose[:3], [np.cos(a2[3] / 2)], np.sin(a2[3] / 2) * a2[:3]])
            fr.dt = None if k == 0 else (k * seq.dt - (k - 1) * seq.dt)
            ep, ev = orc.step(fr)
            assert rel(vel_log[k], ev) < 1e-4 or np.linalg.norm(vel_log[k] - ev) < 1e-9, (k, t)
            assert rel(pose_log[k, :9], ep[:9]) < 1e-4, (k, t)
            eaa = dataset_io.quat_to_axis_angle(ep[9:])
            q_log = np.concatenate([[np.cos(pose_log[k, 12] / 2)], np.sin(pose_log[k, 12] / 2) * pose_log[k, 9:12]])
            q_exp = np.concatenate([[np.cos(eaa[3] / 2)], np.sin(eaa[3] / 2) * eaa[:3]])
            assert quat_close(q_log, q_exp) < 1e-4, (k, t)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,stride", [("f32", 1), ("s16", 4)])
def test_reference_call_sequence_over_adapters_matches_fused_step(hostlib, tmp_path, fmt, stride):
    """The fine-grained boundary: ROFTFilter::filtering_step (ROFTFilter.cpp:255-367) transcribed over the adapter classes
    that carry the reference's names (ROFT::ImageOpticalFlowMeasurement<T>, SKFCorrection, UKFCorrection,
    ImageSegmentationOFAidedSource<T>, CartesianQuaternionMeasurement ...) gives the same beliefs, frame by frame, as the
    fused batched loop (roftb_filter_step) - tests/cpp/adapter_check.cpp."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = small_cfg(flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0, subsampling_radius=float(stride),
                    segm_delay=3, pose_delay=3, sample_time=1.0 / 30.0)
    F = 14
    seq = sequence(cfg, 1, F, flow_format=fmt, target_coverage=0.3)
    root = str(tmp_path / "seq0")
    dataset_io.write_sequence(root, seq, 0, fx=cfg.fx, fy=cfg.fy, cx=cfg.cx, cy=cfg.cy)
    out = subprocess.run([os.path.join(HOST, "adapter_check"), "--sequence", root, "--stride", str(stride), "--desired-fps", "10"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"adapter_check: {F} frames" in out.stdout, out.stdout


def test_adapter_headers_carry_the_reference_names():
    """The adapter header declares every class / virtual the reference's L2 interfaces name."""
    h = open(os.path.join(HOST, "roft_adapters.h")).read()
    for cls in ("class ImageSegmentationOFAidedSource", "class ImageSegmentationMeasurement", "class ImageOpticalFlowMeasurement",
                "class SKFCorrection : public bfl::GaussianCorrection", "class UKFCorrection : public bfl::GaussianCorrection",
                "class CartesianQuaternionModel : public bfl::StateModel", "class CartesianQuaternionMeasurement : public bfl::MeasurementModel",
                "class SpatialVelocityModel : public bfl::LinearStateModel"):
        assert cls in h, cls
    for virt in ("bool freeze(const bfl::Data& data = bfl::Data()) override", "predictedMeasure(const Eigen::Ref<const Eigen::MatrixXd>&",
                 "getMeasurementMatrix() const override", "getNoiseCovarianceMatrix() const override", "setProperty(const std::string& property) override",
                 "void correctStep(const bfl::GaussianMixture& pred_state, bfl::GaussianMixture& corr_state) override",
                 "enum class MeasurementMode { Standard, RepeatOnlyVelocity, PopBufferedMeasurement }",
                 "enum class FreezeType { OnlyStepSource, ExceptStepSource, Complete }"):
        assert virt in h, virt


def test_flow_queue_oracle_semantics():
    """OpticalFlowQueueHandler.cpp:18-57: bounded window, region = the frames AFTER the stamped one, empty if not found."""
    q = o.OpticalFlowQueue(4)
    for k in range(6):
        q.add_flow(np.full((2, 2, 2), k, np.float32), 0.1 * k)
    assert len(q.buffer) == 4                                   # frames 2..5 kept
    assert [int(f[0, 0, 0]) for f in q.get_buffer_region(0.3)] == [4, 5]
    assert q.get_buffer_region(0.5) == []                       # the newest frame: nothing follows it
    assert q.get_buffer_region(0.1) == []                       # fell out of the window
    assert [int(f[0, 0, 0]) for f in q.get_buffer_region(0.2 + 5e-4)] == [3, 4, 5]   # 1 ms matching tolerance


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["f32", "s16"])
def test_stamped_mask_sync_matches_oracle(hostlib, fmt):
    """f2: ImageSegmentationOFAidedSourceStamped (time-stamp matched masks + flow queue, ...Stamped.hpp:153-318) - the
    adapter over roftb_mask_sync against the numpy/cv2 restatement: masks arrive with irregular latency (1-11 frames,
    i.e. chains longer than ROFTB_MAX_DELAY), one stamp matches nothing, one mask is empty."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = small_cfg(W=160, H=96, flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0, segm_delay=0)
    F = 26
    seq = sequence(cfg, 1, F, flow_format=fmt, target_coverage=0.3)
    H, W = cfg.height, cfg.width
    stamps = 100.0 + np.arange(F) / 30.0
    # frame at which a mask computed on frame `src` is delivered
    deliveries = {0: 0, 3: 2, 9: 8, 20: 9, 22: 21, 24: 5}
    masks = np.zeros((F, H, W), np.uint8); mvalid = np.zeros(F, np.uint8); mstamp = np.full(F, -1.0)
    for at, src in deliveries.items():
        masks[at] = seq.mask[src, 0].numpy(); mvalid[at] = 1; mstamp[at] = stamps[src]
    masks[22] = 0                      # an empty (uninformative) mask: skipped
    mstamp[24] = 55.5                  # a stamp that matches no queued frame: falls back to the current flow
    flows = np.ascontiguousarray(seq.flow[:, 0].numpy())
    fvalid = np.ones(F, np.uint8); fvalid[0] = 0
    out = np.zeros((F, H, W), np.uint8); avail = np.zeros(F, np.uint8)
    fn = hostlib.rofth_stamped_sync_run
    fn.argtypes = [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    rc = fn(W, H, 13 if fmt == "f32" else 11, cfg.flow_grid, cfg.flow_scale, F, masks.ctypes.data, mvalid.ctypes.data, mstamp.ctypes.data,
            flows.ctypes.data, fvalid.ctypes.data, stamps.ctypes.data, -1, out.ctypes.data, avail.ctypes.data)
    assert rc == 0
    src = o.StampedOFAidedSegmentationSource(cfg)
    longest = 0
    for k in range(F):
        if mvalid[k]:
            longest = max(longest, len(src.queue.get_buffer_region(mstamp[k])))
        src.step_frame(masks[k] if mvalid[k] else None, mstamp[k], flows[k] if fvalid[k] else None, stamps[k])
        ok, m = src.segmentation()
        assert bool(avail[k]) == ok, k
        if ok:
            assert np.array_equal(out[k], m), (k, int((out[k] != m).sum()))
    assert longest > 8                 # a chain longer than the filter loop's ROFTB_MAX_DELAY went through the operator


CFG_TEXT = """
# flattened and overridden like the reference's ConfigParser
sample_time = 0.033333333333;
camera_dataset:
{
    width = 320; height = 180;      // two settings on one line
    fx = 307.35714; fy = 307.35714; cx = 160.0; cy = 90.0;
    path = "?";
    heading_zeros = 0; index_offset = 0;
}
/* nested groups */
initial_condition: { pose: { v = [0.0, 0.0, 0.0]; cov_q = [0.001, 0.001, 0.001]; axis_angle = [1.0, 0.0, 0.0, 0.0]; } }
measurement_model = { velocity: { subsampling_radius = 35.0; weight_flow = true; cov_flow = [1.0, 1.0]; } use_pose = true; }
segmentation_dataset: { set = "mrcnn"; desired_fps = 5.0; counts = [1, 2, 3]; }
"""


def _config_dump(hostlib, text, overrides=()):
    buf = ctypes.create_string_buffer(1 << 16)
    paths = (ctypes.c_char_p * max(1, len(overrides)))(*[p.encode() for p, _ in overrides])
    vals = (ctypes.c_char_p * max(1, len(overrides)))(*[v.encode() for _, v in overrides])
    rc = hostlib.rofth_config_dump(text.encode(), paths, vals, len(overrides), buf, len(buf))
    out = buf.value.decode()
    if rc != 0:
        return rc, out
    d = {}
    for line in out.strip().split("\n"):
        k, t, v = line.split("\t")
        d[k] = (t, v)
    return 0, d


def test_config_parser_grammar_and_overrides(hostlib):
    """f3: the libconfig subset + `--a::b::c value` overrides of the reference's front-end (ConfigParser.cpp:8-169)."""
    rc, d = _config_dump(hostlib, CFG_TEXT)
    assert rc == 0
    assert d["sample_time"] == ("float", "0.033333333333")
    assert d["camera_dataset.width"] == ("int", "320") and d["camera_dataset.path"] == ("string", "?")
    assert d["initial_condition.pose.axis_angle"] == ("array", "1.0,0.0,0.0,0.0")
    assert d["measurement_model.velocity.weight_flow"] == ("bool", "true") and d["measurement_model.use_pose"] == ("bool", "true")
    assert d["segmentation_dataset.counts"] == ("array", "1,2,3")
    rc, d = _config_dump(hostlib, CFG_TEXT, [("measurement_model::velocity::subsampling_radius", "1.0"),
                                             ("measurement_model::velocity::weight_flow", "false"),
                                             ("measurement_model::velocity::cov_flow", "2.0, 0.5"),
                                             ("segmentation_dataset::set", "gt"), ("camera_dataset::width", "640")])
    assert rc == 0
    assert d["measurement_model.velocity.subsampling_radius"][1] == "1.0" and d["measurement_model.velocity.weight_flow"][1] == "false"
    assert d["measurement_model.velocity.cov_flow"][1] == "2.0,0.5" and d["segmentation_dataset.set"][1] == "gt"
    assert d["camera_dataset.width"][1] == "640"
    # errors: unknown option, wrong type, wrong array length, syntax error with file:line
    assert _config_dump(hostlib, CFG_TEXT, [("camera_dataset::nope", "1")])[0] == -1
    assert _config_dump(hostlib, CFG_TEXT, [("measurement_model::velocity::weight_flow", "maybe")])[0] == -1
    assert _config_dump(hostlib, CFG_TEXT, [("measurement_model::velocity::cov_flow", "1.0")])[0] == -1
    rc, msg = _config_dump(hostlib, "a = 1;\nb: { c = ; }\n")
    assert rc == -1 and "<string>:2" in msg


def test_config_parser_reads_the_reference_files(hostlib):
    """The reference's own configuration files parse, with the values SURVEY.md quotes (skipped where the read-only
    reference tree is not mounted, e.g. on the GPU box)."""
    path = "/root/reference/config/config_fast_ycb.cfg"
    if not os.path.exists(path):
        pytest.skip("reference tree not available")
    rc, d = _config_dump(hostlib, open(path).read())
    assert rc == 0, d
    assert d["camera_dataset.width"][1] == "1280" and d["camera_dataset.fx"][1] == "1229.4285612615463"
    assert d["measurement_model.velocity.subsampling_radius"][1] == "35.0" and d["measurement_model.velocity.weight_flow"][1] == "true"
    assert d["segmentation_dataset.desired_fps"][1] == "5.0" and d["pose_dataset.delay"][1] == "true"
    assert d["unscented_transform.alpha"][1] == "1.0" and d["outlier_rejection.gain"][1] == "0.01"
    rc, d2 = _config_dump(hostlib, open("/root/reference/config/config_ho3d.cfg").read())
    assert rc == 0 and d2["camera_dataset.width"][1] == "640"


@pytest.mark.gpu
def test_tracker_executable_config_front_end(hostlib, tmp_path):
    """`roft_b200_tracker --from file.cfg --group::key value ...` (main.cpp:41-424) gives the same log as the flag front-end."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = small_cfg(subsampling_radius=4.0, segm_delay=3, pose_delay=3, sample_time=1.0 / 30.0)
    F = 8
    seq = sequence(cfg, 1, F, target_coverage=0.3)
    root = str(tmp_path / "seq0")
    dataset_io.write_sequence(root, seq, 0, fx=cfg.fx, fy=cfg.fy, cx=cfg.cx, cy=cfg.cy)
    aa = [float(v) for v in dataset_io.quat_to_axis_angle(seq.pose[0, 0].numpy()[3:])]
    x0 = [float(v) for v in seq.pose[0, 0, :3].numpy()]
    text = f"""
sample_time = 0.033333333333;
camera_dataset: {{ width = {cfg.width}; height = {cfg.height}; fx = {cfg.fx!r}; fy = {cfg.fy!r}; cx = {cfg.cx!r}; cy = {cfg.cy!r};
                  path = "?"; heading_zeros = 0; index_offset = 0; }}
initial_condition: {{
  pose: {{ v = [0.0, 0.0, 0.0]; w = [0.0, 0.0, 0.0]; x = [{x0[0]!r}, {x0[1]!r}, {x0[2]!r}]; axis_angle = [{aa[0]!r}, {aa[1]!r}, {aa[2]!r}, {aa[3]!r}];
          cov_v = [0.001, 0.001, 0.001]; cov_w = [0.001, 0.001, 0.001]; cov_x = [0.001, 0.001, 0.001]; cov_q = [0.001, 0.001, 0.001]; }}
  velocity: {{ v = [0.0, 0.0, 0.0]; w = [0.0, 0.0, 0.0]; cov_v = [0.001, 0.001, 0.001]; cov_w = [0.001, 0.001, 0.001]; }} }}
kinematic_model: {{ pose: {{ sigma_linear = [1.0, 1.0, 1.0]; sigma_angular = [1.0, 1.0, 1.0]; }}
                   velocity: {{ sigma_linear = [0.1, 0.1, 0.1]; sigma_angular = [0.1, 0.1, 0.1]; }} }}
log: {{ enable = true; enable_segmentation = false; path = "?"; }}
measurement_model: {{ pose: {{ cov_v = [0.1, 0.1, 0.1]; cov_w = [0.0001, 0.0001, 0.0001]; cov_x = [0.001, 0.001, 0.001]; cov_q = [0.0001, 0.0001, 0.0001]; }}
                     velocity: {{ cov_flow = [1.0, 1.0]; depth_maximum = 2.0; subsampling_radius = 35.0; weight_flow = true; }}
                     use_pose = true; use_pose_resync = true; use_velocity = true; }}
model: {{ name = "003_cracker_box"; }}
optical_flow_dataset: {{ path = "?"; set = "nvof"; heading_zeros = 0; index_offset = 0; }}
outlier_rejection: {{ enable = false; gain = 0.01; }}
pose_dataset: {{ path = "?"; skip_rows = 0; skip_cols = 0; fps_reduction = true; delay = true; original_fps = 30.0; desired_fps = 5.0; }}
segmentation_dataset: {{ path = "?"; format = "pgm"; set = "gt"; heading_zeros = 0; index_offset = 0; fps_reduction = true; delay = true;
                        original_fps = 30.0; desired_fps = 5.0; flow_aided = true; }}
unscented_transform: {{ alpha = 1.0; beta = 2.0; kappa = 0.0; }}
"""
    cfg_path = tmp_path / "tracker.cfg"
    cfg_path.write_text(text)
    log_a, log_b = tmp_path / "log_a", tmp_path / "log_b"
    log_a.mkdir(); log_b.mkdir()
    exe = os.path.join(HOST, "roft_b200_tracker")
    a = subprocess.run([exe, "--from", str(cfg_path), "--camera_dataset::path", root, "--optical_flow_dataset::path", root,
                        "--segmentation_dataset::path", root, "--pose_dataset::path", root + "/gt/poses.txt", "--log::path", str(log_a),
                        "--measurement_model::velocity::subsampling_radius", "4.0", "--segmentation_dataset::desired_fps", "10.0",
                        "--pose_dataset::desired_fps", "10.0"], capture_output=True, text=True)
    assert a.returncode == 0, a.stdout + a.stderr
    b = subprocess.run([exe, "--sequence", root, "--log", str(log_b), "--stride", "4", "--desired-fps", "10"], capture_output=True, text=True)
    assert b.returncode == 0, b.stdout + b.stderr
    pa = np.loadtxt(log_a / "pose_estimate.txt"); pb = np.loadtxt(log_b / "pose_estimate.txt")
    va = np.loadtxt(log_a / "velocity_estimate.txt"); vb = np.loadtxt(log_b / "velocity_estimate.txt")
    assert pa.shape == (F, 13) and np.allclose(pa, pb, rtol=1e-9, atol=1e-12) and np.allclose(va, vb, rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_tracker_executable_with_outlier_rejection(hostlib, tmp_path):
    """roft_b200_tracker --outlier-rejection --mesh box.obj (ROFTFilter ctor parameters pose_outlier_rejection + model,
    ROFTFilter.cpp:52-54, 184-199): a grossly displaced pose delivery is rejected like the oracle rejects it."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from roft_b200.synthetic import cuboid_mesh
    cfg = small_cfg(subsampling_radius=4.0, segm_delay=3, pose_delay=3, sample_time=1.0 / 30.0, outlier_rejection=True)
    F = 14
    seq = sequence(cfg, 1, F, target_coverage=0.3, corrupt=False)
    seq.pose[6, 0, :3] += torch.tensor([0.0, 0.07, 0.06], dtype=seq.pose.dtype)
    root = str(tmp_path / "seq0")
    dataset_io.write_sequence(root, seq, 0, fx=cfg.fx, fy=cfg.fy, cx=cfg.cx, cy=cfg.cy, mask_format="png")  # masks as PNG, like Fast-YCB
    verts, faces = cuboid_mesh(seq.half[0].numpy())
    dataset_io.write_obj(str(tmp_path / "box.obj"), verts, faces)
    out = subprocess.run([os.path.join(HOST, "roft_b200_tracker"), "--sequence", root, "--log", str(tmp_path), "--stride", "4",
                          "--desired-fps", "10", "--outlier-rejection", "--mesh", str(tmp_path / "box.obj"), "--mask-format", "png"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    pose_log = np.loadtxt(tmp_path / "pose_estimate.txt")
    x0 = np.zeros(13)
    aa = dataset_io.quat_to_axis_angle(seq.pose[0, 0].numpy()[3:])
    x0[6:9] = seq.pose[0, 0, :3].numpy()
    x0[9] = np.cos(aa[3] / 2); x0[10:] = np.sin(aa[3] / 2) * aa[:3]
    orc = o.RoftFilterOracle(cfg, x0, mesh=(verts, faces))
    for k in range(F):
        fr = frame_inputs(seq, cfg, k, 0)
        if fr.pose is not None:
            a2 = dataset_io.quat_to_axis_angle(fr.pose[3:])
            fr.pose = np.concatenate([fr.pose[:3], [np.cos(a2[3] / 2)], np.sin(a2[3] / 2) * a2[:3]])
        fr.dt = None if k == 0 else (k * seq.dt - (k - 1) * seq.dt)
        ep, _ = orc.step(fr)
        assert rel(pose_log[k, :9], ep[:9]) < 1e-4, (k, orc.or_selected)
    assert 1 in orc.or_selected and 0 in orc.or_selected, orc.or_selected
    # a missing mesh is an error, not a silent fallback
    bad = subprocess.run([os.path.join(HOST, "roft_b200_tracker"), "--sequence", root, "--log", str(tmp_path), "--outlier-rejection",
                          "--mesh", str(tmp_path / "nope.obj"), "--mask-format", "png"], capture_output=True, text=True)
    assert bad.returncode != 0 and "cannot open" in (bad.stdout + bad.stderr)


def test_png_mask_reader_matches_opencv(hostlib, tmp_path):
    """read_png_gray8 = cv::imread(IMREAD_UNCHANGED) + convertTo(CV_8UC1) (DatasetImageSegmentation.cpp:130-132) on 8- and
    16-bit greyscale PNGs written by OpenCV at several compression levels (all five scanline filters occur)."""
    import cv2
    rng = np.random.default_rng(7)
    hostlib.rofth_read_png.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p]
    cases = []
    m = np.zeros((90, 160), np.uint8); m[20:70, 40:120] = 255; cases.append(m)                     # a Mask R-CNN style mask
    cases.append(rng.integers(0, 256, (33, 47)).astype(np.uint8))                                   # noise, odd size
    g = (np.add.outer(np.arange(64), np.arange(96)) * 2 % 256).astype(np.uint8); cases.append(g)   # gradient (sub / up / paeth filters)
    cases.append((rng.integers(0, 600, (40, 50))).astype(np.uint16))                                # 16 bit: saturates to 255
    for k, img in enumerate(cases):
        for level in (0, 3, 9):
            p = str(tmp_path / f"m{k}_{level}.png")
            assert cv2.imwrite(p, img, [cv2.IMWRITE_PNG_COMPRESSION, level])
            exp = cv2.imread(p, cv2.IMREAD_UNCHANGED)
            exp = np.clip(exp, 0, 255).astype(np.uint8)
            out = np.zeros(img.size, np.uint8)
            c = ctypes.c_ulonglong(0); r = ctypes.c_ulonglong(0)
            assert hostlib.rofth_read_png(p.encode(), out.ctypes.data, out.size, ctypes.byref(c), ctypes.byref(r)) == 0
            assert (r.value, c.value) == img.shape and np.array_equal(out.reshape(img.shape), exp)
    # colour PNGs / garbage are refused, not misread
    p = str(tmp_path / "rgb.png")
    cv2.imwrite(p, rng.integers(0, 256, (8, 8, 3)).astype(np.uint8))
    out = np.zeros(64 * 3, np.uint8); c = ctypes.c_ulonglong(0); r = ctypes.c_ulonglong(0)
    assert hostlib.rofth_read_png(p.encode(), out.ctypes.data, out.size, ctypes.byref(c), ctypes.byref(r)) == -1
    (tmp_path / "bad.png").write_bytes(b"not a png at all")
    assert hostlib.rofth_read_png(str(tmp_path / "bad.png").encode(), out.ctypes.data, out.size, ctypes.byref(c), ctypes.byref(r)) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("resync", [True, False])
def test_reference_call_sequence_with_outlier_rejection(hostlib, tmp_path, resync):
    """adapter_check --mesh: ROFTFilter::filtering_step INCLUDING correct_outlier_rejection / pick_best_alternative /
    buffer_outlier_rejection_features (ROFTFilter.cpp:313-359, 467-676) transcribed over the adapter classes and the
    render / pick-best operators, against the fused device-side path of roftb_filter_step (cfg.outlier_rejection)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from roft_b200.synthetic import cuboid_mesh
    cfg = small_cfg(subsampling_radius=4.0, segm_delay=3, pose_delay=3, sample_time=1.0 / 30.0)
    F = 20
    seq = sequence(cfg, 1, F, corrupt=False)
    seq.pose[6, 0, :3] += torch.tensor([0.0, 0.07, 0.06], dtype=seq.pose.dtype)
    seq.pose[12, 0, :3] += torch.tensor([0.09, 0.0, 0.0], dtype=seq.pose.dtype)
    root = str(tmp_path / "seq0")
    dataset_io.write_sequence(root, seq, 0, fx=cfg.fx, fy=cfg.fy, cx=cfg.cx, cy=cfg.cy)
    dataset_io.write_obj(str(tmp_path / "box.obj"), *cuboid_mesh(seq.half[0].numpy()))
    args = [os.path.join(HOST, "adapter_check"), "--sequence", root, "--stride", "4", "--desired-fps", "10", "--mesh", str(tmp_path / "box.obj")]
    out = subprocess.run(args + ([] if resync else ["--no-resync"]), capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"adapter_check: {F} frames" in out.stdout, out.stdout
    choices = out.stdout.split("outlier rejection choices:")[1].split()
    assert "0" in choices and "1" in choices, out.stdout
