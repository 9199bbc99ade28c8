"""roft_b200/metrics.py (evaluation/metrics.py restated) against golden vectors produced by the reference's own
tools/third_party/bop_pose_error.py (tests/golden/make_metrics_golden.py)."""
import os

import numpy as np

from roft_b200 import metrics as m

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz"))


def test_add_adi_and_auc_match_the_reference_module():
    ref, sig, pts = G["reference"], G["signal"], G["points"]
    d_add, auc_add = m.auc(ref, sig, pts, "add")
    d_adi, auc_adi = m.auc(ref, sig, pts, "adi")
    assert np.allclose(d_add, G["add"], rtol=1e-12, atol=0)
    assert np.allclose(d_adi, G["adi"], rtol=1e-12, atol=0)
    assert abs(auc_add - float(G["auc_add"])) < 1e-9 and abs(auc_adi - float(G["auc_adi"])) < 1e-9
    assert (d_add > 0.1).any() and (d_add < 0.02).any()          # the fixture spans both sides of the 10 cm threshold
    assert np.all(d_adi <= d_add + 1e-12)                        # nearest-neighbour distance never exceeds the paired one


def test_vocap_corner_cases():
    assert m.vocap(np.array([np.inf, np.inf]), np.array([0.5, 1.0])) == float(G["vocap_empty"]) == 0.0
    v = m.vocap(np.array([0.01, 0.01, 0.05, np.inf]), np.array([0.25, 0.5, 0.75, 1.0], np.float32))
    assert abs(v - float(G["vocap_ties"])) < 1e-12


def test_rmse_and_time_metrics():
    rng = np.random.default_rng(0)
    ref = np.zeros((50, 7)); ref[:, :3] = rng.normal(size=(50, 3)); ref[:, 3:6] = [0, 0, 1]; ref[:, 6] = rng.uniform(0, 1, 50)
    sig = ref.copy(); sig[:, 0] += 0.02; sig[:, 6] += np.radians(3.0)
    assert abs(m.rmse_cartesian_3d(ref, sig) - 2.0) < 1e-12            # 2 cm
    assert abs(m.rmse_angular(ref, sig) - 3.0) < 1e-9                  # 3 degrees about the same axis
    v_ref = rng.normal(size=(50, 6)); v_sig = v_ref.copy(); v_sig[:, 1] += 0.05; v_sig[:, 4] += np.radians(2.0)
    assert abs(m.rmse_linear_velocity(v_ref, v_sig) - 5.0) < 1e-12     # 5 cm/s
    assert abs(m.rmse_angular_velocity(v_ref, v_sig) - 2.0) < 1e-12    # 2 deg/s
    t = np.array([[10.0, 1.0], [40.0, 2.0], [33.0, 0.0], [35.5, 0.0]])
    assert m.mean_time(t) == 29.625 and m.time_excess_33_ms(t) == 2.0  # strictly more than 33 ms
    # (x, q) -> (x, axis, angle) as the filter logs it
    q = np.array([[0.1, 0.2, 0.3, np.cos(0.4), 0.0, np.sin(0.4), 0.0], [0, 0, 0, -np.cos(0.4), 0.0, np.sin(0.4), 0.0]])
    aa = m.quat_pose_to_axis_angle(q)
    assert np.allclose(aa[0], [0.1, 0.2, 0.3, 0, 1, 0, 0.8]) and np.allclose(aa[1, 3:], [0, -1, 0, 0.8])
