"""Generates tests/golden/*.npz from the CPU oracle (oracle/roft_oracle.py, whose mask path runs the genuine OpenCV
primitives).  The reference ships no golden vectors and cannot be built or imported in this image (C++ over Eigen /
OpenCV / bfl / RobotsIO, all absent) - so these fixtures pin the ORACLE (regression) and give the GPU tests fixed,
seed-independent inputs; they do not pin the reference itself ("parity unpinned").

    python tests/golden/make_golden.py        # rewrites the fixtures next to this file
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import roft_oracle as o  # noqa: E402
from helpers import frame_inputs, sequence, small_cfg  # noqa: E402

W, H = 96, 64


def cfg_of(fmt, **kw):
    return small_cfg(W, H, flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0, **kw)


def main():
    rng = np.random.default_rng(20221017)
    # ---- mask synchronisation: mixed values, collisions, NaN/inf flows, Q2 origin ---------------------------------
    for fmt in ("f32", "s16"):
        cfg = cfg_of(fmt, segm_delay=3)
        mask = rng.choice(np.array([0, 1, 2, 255], np.uint8), size=(H, W), p=[0.55, 0.1, 0.15, 0.2])
        mask[0, 0] = 255
        flows = []
        for _ in range(3):
            if fmt == "f32":
                f = rng.normal(0, 2.0, (H, W, 2)).astype(np.float32)
                bad = rng.random((H, W))
                f[bad < 0.02] = np.nan
                f[(bad > 0.02) & (bad < 0.03)] = np.inf
                f[(bad > 0.03) & (bad < 0.04)] = 3e9
            else:
                f = rng.integers(-120, 120, (H // 4, W // 4, 2)).astype(np.int16)
            flows.append(f)
        out = {}
        for name, zo, fl in (("new", False, flows), ("prop", True, flows[-1:])):
            m = mask.copy()
            if zo:
                m[0, 0] = 0
            raw = o.remap_exact(m, o.mask_warp_map(m, fl, cfg))
            out[f"raw_{name}"] = raw
            out[f"thr_{name}"] = o.threshold_mask(raw)
        np.savez_compressed(os.path.join(HERE, f"mask_sync_{fmt}.npz"), mask=mask, flows=np.stack(flows), **out)

    # ---- velocity measurement + sequential SKF, UKF, filter loop on a synthetic sequence ---------------------------
    for fmt in ("f32", "s16"):
        cfg = cfg_of(fmt, subsampling_radius=2.0, segm_delay=3, pose_delay=3)
        seq = sequence(cfg, 1, 10, flow_format=fmt, target_coverage=0.3, seed=77)
        m = o.threshold_mask(seq.mask[1, 0].numpy()); d = seq.depth[1, 0].numpy(); f = seq.flow[2, 0].numpy()
        z, Hm, _ = o.flow_velocity_measurement(m, d, f, cfg, cfg.sample_time)
        xp = np.array([0.02, -0.1, 0.03, -0.3, 0.2, 0.1]); Pp = np.eye(6) * 0.101
        xs, Ps = o.skf_correct(xp, Pp, z, Hm, np.diag(cfg.cov_flow), True)
        _, _, Lm, em = o.skf_correct_information(xp, Pp, z, Hm, np.diag(cfg.cov_flow), True)
        x0 = np.zeros(13); x0[6:] = seq.pose[0, 0].numpy()
        orc = o.RoftFilterOracle(cfg, x0)
        pms, vms, raws = [], [], []
        for k in range(10):
            ep, ev = orc.step(frame_inputs(seq, cfg, k, 0))
            pms.append(ep); vms.append(ev); raws.append(orc.seg_source.mask.copy())
        np.savez_compressed(os.path.join(HERE, f"sequence_{fmt}.npz"), depth=seq.depth[:, 0].numpy(), flow=seq.flow[:, 0].numpy(),
                            mask=seq.mask[:, 0].numpy(), pose=seq.pose[:, 0].numpy(), pose_valid=seq.pose_valid[:, 0].numpy(),
                            z=z, H=Hm, x_pred=xp, P_pred=Pp, x_seq=xs, P_seq=Ps, lam=Lm, eta=em,
                            p_mean=np.stack(pms), v_mean=np.stack(vms), raw=np.stack(raws))
    # ---- UKF predict / correct -----------------------------------------------------------------------------------
    cfg = cfg_of("f32")
    mean = np.zeros(13); mean[:9] = rng.normal(0, 0.3, 9); mean[8] += 0.7
    q = rng.normal(size=4); mean[9:] = q / np.linalg.norm(q)
    B = rng.normal(size=(12, 12)) * 0.02
    cov = B @ B.T + np.diag(rng.uniform(1e-4, 2e-3, 12))
    pm, pc = o.ukf_predict(mean, cov, cfg, 0.0333)
    meas = np.zeros(13)
    meas[:6] = pm[:6] + rng.normal(0, 0.05, 6)
    meas[6:9] = pm[6:9] + rng.normal(0, 0.01, 3)
    meas[9:] = o.sum_quaternion_rotation_vector(pm[9:], rng.normal(0, 0.05, 3))[0]
    out = dict(mean=mean, cov=cov, pred_mean=pm, pred_cov=pc, meas=meas)
    for name, mt, mv in (("vel", o.MEAS_VELOCITY, meas[:6]), ("pose", o.MEAS_POSE, meas[6:]), ("pv", o.MEAS_POSE_VELOCITY, meas)):
        cm, cc = o.ukf_correct(pm, pc, mv, mt, cfg)
        out[f"corr_mean_{name}"] = cm; out[f"corr_cov_{name}"] = cc
    np.savez_compressed(os.path.join(HERE, "ukf.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
