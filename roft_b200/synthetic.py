"""Deterministic synthetic Fast-YCB-format sequences.

Scene: a rigid cuboid with the 003_cracker_box extents moving with a constant twist in front
of a pinhole camera; depth by ray/cuboid intersection, optical flow by re-projecting the moved
surface point, mask = silhouette.  Gate-exercising corruption is injected exactly as the survey
prescribes (depth holes / out-of-range, NaN and 1e10 flow entries).

Written with torch ops so the same code generates the small CPU test sequences and the
multi-GB device-resident benchmark sequences (inputs for a throughput run cannot come over
PCIe).  torch is plumbing here (device memory + RNG); nothing on the tracked path uses it.

Frame conventions: flow[k] is the displacement of pixels from frame k-1 to
frame k (flow[0] does not exist); depth in metres; mask 255/0.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

CRACKER_BOX_EXTENTS = (0.164, 0.213, 0.072)  # metres, measured from meshes/DOPE/003_cracker_box.obj
BASE_SEED = 0x20F7


@dataclass
class SyntheticSequence:
    depth: torch.Tensor        # [F, T, H, W] float32
    flow: torch.Tensor         # [F, T, Hf, Wf, 2] float32 (grid 1) or int16 (grid 4, S10.5); flow[0] is zeros/unused
    mask: torch.Tensor         # [F, T, H, W] uint8 ground-truth silhouette at each frame
    pose: torch.Tensor         # [F, T, 7] float64 noisy pose measurement (x, q wxyz) of each frame
    pose_valid: torch.Tensor   # [F, T] bool (False = all-zero row in poses.txt)
    gt_pose: torch.Tensor      # [F, T, 7] float64
    gt_twist: torch.Tensor     # [T, 6] float64 (v of the object origin, w), camera frame
    flow_grid: int
    flow_scale: float
    dt: float
    half: Optional[torch.Tensor] = None   # [T, 3] float64 half extents of each track's cuboid


def cuboid_mesh(half):
    """Triangle mesh (vertices [8,3] float32, faces [12,3] int32) of the axis-aligned cuboid with the given half extents -
    the object every synthetic track shows, for the render-and-compare pose test (roftb_set_mesh)."""
    import numpy as np
    h = np.asarray(half, np.float64)
    v = np.array([[sx * h[0], sy * h[1], sz * h[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32)
    f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4],
                  [1, 5, 7], [1, 7, 3]], np.int32)
    return v, f


def _quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    w, x, y, z = q.unbind(-1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(q.shape[:-1] + (3, 3))


def _quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw], -1)


def _rotvec_to_quat(r: torch.Tensor) -> torch.Tensor:
    n = r.norm(dim=-1, keepdim=True)
    half = 0.5 * n
    k = torch.where(n > 1e-12, torch.sin(half) / n.clamp_min(1e-12), torch.full_like(n, 0.5))
    return torch.cat([torch.cos(half), k * r], -1)


def _render(x, R, half, K, H, W, background):
    """Ray / cuboid slab intersection. x [T,3], R [T,3,3], half [T,3] -> depth [T,H,W], hit [T,H,W]."""
    fx, fy, cx, cy = K
    dev = x.device
    u = (torch.arange(W, device=dev, dtype=torch.float32) - cx) / fx
    v = (torch.arange(H, device=dev, dtype=torch.float32) - cy) / fy
    # ray dir in camera frame (u, v, 1); object frame: d' = R^T d, o' = -R^T x
    Rt = R.transpose(1, 2).to(torch.float32)
    o = -(Rt @ x.to(torch.float32).unsqueeze(-1)).squeeze(-1)  # [T,3]
    d = (Rt[:, :, 0].reshape(-1, 3, 1, 1) * u.reshape(1, 1, 1, W)
         + Rt[:, :, 1].reshape(-1, 3, 1, 1) * v.reshape(1, 1, H, 1)
         + Rt[:, :, 2].reshape(-1, 3, 1, 1))  # [T,3,H,W]
    inv = 1.0 / torch.where(d.abs() < 1e-9, torch.full_like(d, 1e-9), d)
    hb = half.to(torch.float32).reshape(-1, 3, 1, 1)
    ob = o.reshape(-1, 3, 1, 1)
    t1 = (-hb - ob) * inv
    t2 = (hb - ob) * inv
    tn = torch.minimum(t1, t2).amax(dim=1)
    tf = torch.maximum(t1, t2).amin(dim=1)
    hit = (tn <= tf) & (tn > 0.05)
    depth = torch.where(hit, tn, torch.full_like(tn, background))
    return depth, hit


def make_sequence(n_tracks: int, n_frames: int, width: int = 1280, height: int = 720,
                  fx: float = 1229.4285612615463, fy: float = 1229.4285612615463,
                  cx: float = 640.0, cy: float = 360.0, dt: float = 1.0 / 30.0,
                  device: str = "cpu", seed: int = BASE_SEED, flow_format: str = "f32",
                  target_coverage: Optional[float] = None, corrupt: bool = True,
                  flow_noise: float = 0.25, track_chunk: int = 16,
                  mixed_mask_values: bool = False, first_track_id: int = 0) -> SyntheticSequence:
    """Generate ``n_frames`` frames for ``n_tracks`` independent object tracks.

    flow_format: "f32" (CV_32FC2, grid 1, scale 1) or "s16" (CV_16SC2, grid 4, scale 32 - NVOF1).
    target_coverage: if set, the cuboid is rescaled so the silhouette covers about that fraction.
    mixed_mask_values: paint part of each mask with values {1, 2} (pins winner resolution and the
                       ``> 1`` threshold); otherwise masks are 255/0 like Mask R-CNN output.
    Seeds are ``seed + track_id`` so a track's data does not depend on how tracks are partitioned
    across GPUs.
    """
    dev = torch.device(device)
    T, F, H, W = n_tracks, n_frames, height, width
    K = (fx, fy, cx, cy)
    grid = 1 if flow_format == "f32" else 4
    scale = 1.0 if flow_format == "f32" else 32.0
    Hf, Wf = H // grid, W // grid

    depth = torch.empty((F, T, H, W), dtype=torch.float32, device=dev)
    mask = torch.empty((F, T, H, W), dtype=torch.uint8, device=dev)
    flow = torch.zeros((F, T, Hf, Wf, 2), dtype=torch.float32 if grid == 1 else torch.int16, device=dev)
    gt_pose = torch.empty((F, T, 7), dtype=torch.float64)
    pose = torch.empty((F, T, 7), dtype=torch.float64)
    pose_valid = torch.ones((F, T), dtype=torch.bool)
    gt_twist = torch.empty((T, 6), dtype=torch.float64)

    # per-track scene parameters (CPU generators: identical on every device)
    x0 = torch.empty((T, 3), dtype=torch.float64)
    q0 = torch.empty((T, 4), dtype=torch.float64)
    half = torch.empty((T, 3), dtype=torch.float64)
    for t in range(T):
        g = torch.Generator().manual_seed(seed + first_track_id + t)
        r = torch.rand(16, generator=g, dtype=torch.float64)
        x0[t] = torch.tensor([-0.1 + 0.2 * r[0], -0.1 + 0.2 * r[1], 0.45 + 0.45 * r[2]])
        qq = torch.randn(4, generator=g, dtype=torch.float64)
        q0[t] = qq / qq.norm()
        vdir = torch.randn(3, generator=g, dtype=torch.float64)
        wdir = torch.randn(3, generator=g, dtype=torch.float64)
        gt_twist[t, :3] = vdir / vdir.norm() * (0.4 * r[3])
        gt_twist[t, 3:] = wdir / wdir.norm() * (2.0 * r[4])
        half[t] = torch.tensor(CRACKER_BOX_EXTENTS, dtype=torch.float64) * 0.5
        png = torch.randn((F, 6), generator=g, dtype=torch.float64)
        inval = torch.rand(F, generator=g) < 0.03
        for f in range(F):
            tt = f * dt
            xq = _quat_mul(_rotvec_to_quat(gt_twist[t, 3:] * tt), q0[t])
            gt_pose[f, t, :3] = x0[t] + gt_twist[t, :3] * tt
            gt_pose[f, t, 3:] = xq
            # noisy pose measurement: N(0, 2 mm), N(0, 1 deg)
            dq = _rotvec_to_quat(png[f, 3:] * math.radians(1.0))
            pose[f, t, :3] = gt_pose[f, t, :3] + png[f, :3] * 0.002
            pose[f, t, 3:] = _quat_mul(dq, xq)
            if corrupt and inval[f] and f > 0:
                pose_valid[f, t] = False
                pose[f, t] = 0.0

    # one noise generator per GLOBAL track id: a track's images do not depend on the partition / chunking
    gens = [torch.Generator(device=dev).manual_seed(seed * 7919 + first_track_id + t) for t in range(T)]

    def randn(c0, c1, shape):
        return torch.stack([torch.randn(shape, generator=gens[t], device=dev, dtype=torch.float32) for t in range(c0, c1)])

    def rand(c0, c1, shape):
        return torch.stack([torch.rand(shape, generator=gens[t], device=dev, dtype=torch.float32) for t in range(c0, c1)])

    half_used = half.clone()
    for c0 in range(0, T, track_chunk):
        c1 = min(T, c0 + track_chunk)
        hc = half[c0:c1].to(dev)
        if target_coverage is not None:
            # rescale the cuboid so that the frame-0 silhouette covers ~target_coverage of the frame
            for _ in range(3):
                _, hit = _render(gt_pose[0, c0:c1, :3].to(dev), _quat_to_rot(gt_pose[0, c0:c1, 3:].to(dev)), hc, K, H, W, 1.5)
                cov = hit.float().mean(dim=(1, 2)).clamp_min(1e-4).to(torch.float64)
                hc = hc * torch.sqrt(target_coverage / cov).clamp(0.25, 4.0).unsqueeze(-1)
                zmin = gt_pose[0, c0:c1, 2].to(dev) - 0.12
                hc = torch.minimum(hc, zmin.clamp_min(0.05).unsqueeze(-1).expand_as(hc) * torch.tensor([4.0, 4.0, 1.0], device=dev, dtype=torch.float64))
        half_used[c0:c1] = hc.to("cpu", torch.float64)
        prev = None
        for f in range(F):
            xf = gt_pose[f, c0:c1, :3].to(dev)
            Rf = _quat_to_rot(gt_pose[f, c0:c1, 3:].to(dev))
            d, hit = _render(xf, Rf, hc, K, H, W, 1.5)
            m = torch.where(hit, torch.full_like(d, 255), torch.zeros_like(d)).to(torch.uint8)
            if mixed_mask_values:
                band = (torch.arange(W, device=dev) % 7 == 0).reshape(1, 1, W)
                band2 = (torch.arange(H, device=dev) % 5 == 0).reshape(1, H, 1)
                m = torch.where(hit & band, torch.full_like(m, 2), m)
                m = torch.where(hit & band2 & ~band, torch.full_like(m, 1), m)
            mask[f, c0:c1] = m
            if prev is not None:
                # flow k-1 -> k evaluated at frame k-1 pixels
                dp, hitp, xp, Rp = prev
                u = torch.arange(W, device=dev, dtype=torch.float32).reshape(1, 1, W)
                v = torch.arange(H, device=dev, dtype=torch.float32).reshape(1, H, 1)
                X = torch.stack([(u - cx) / fx * dp, (v - cy) / fy * dp, dp], -1)  # [t,H,W,3]
                dR = (Rf @ Rp.transpose(1, 2)).to(torch.float32)
                Xr = X - xp.to(torch.float32).reshape(-1, 1, 1, 3)
                Xn = torch.einsum("tij,thwj->thwi", dR, Xr) + xf.to(torch.float32).reshape(-1, 1, 1, 3)
                un = fx * Xn[..., 0] / Xn[..., 2] + cx
                vn = fy * Xn[..., 1] / Xn[..., 2] + cy
                fl = torch.stack([un - u, vn - v], -1)
                bg = torch.tensor([0.3, -0.2], device=dev, dtype=torch.float32)
                fl = torch.where(hitp.unsqueeze(-1), fl, bg.expand_as(fl))
                if flow_noise > 0:
                    # optical-flow error is spatially smooth: coarse-grid noise, bilinearly upsampled,
                    # plus a small i.i.d. component
                    cg = randn(c0, c1, (2, max(2, H // 16), max(2, W // 16)))
                    sm = torch.nn.functional.interpolate(cg, size=(H, W), mode="bilinear", align_corners=True)
                    fl = fl + flow_noise * sm.permute(0, 2, 3, 1)
                    fl = fl + (0.1 * flow_noise) * randn(c0, c1, tuple(fl.shape[1:]))
                if grid == 1:
                    if corrupt:
                        r = rand(c0, c1, tuple(fl.shape[1:-1]))
                        fl = torch.where((r < 0.005).unsqueeze(-1), torch.full_like(fl, float("nan")), fl)
                        fl = torch.where(((r >= 0.005) & (r < 0.01)).unsqueeze(-1), torch.full_like(fl, 1e10), fl)
                    flow[f, c0:c1] = fl
                else:
                    blk = fl.reshape(-1, Hf, grid, Wf, grid, 2).mean(dim=(2, 4))
                    flow[f, c0:c1] = torch.round(blk * scale).clamp(-32768, 32767).to(torch.int16)
            if corrupt:
                r = rand(c0, c1, tuple(d.shape[1:]))
                dn = torch.where(r < 0.02, torch.zeros_like(d), d)
                dn = torch.where((r >= 0.02) & (r < 0.03), torch.full_like(d, 3.0), dn)
            else:
                dn = d
            depth[f, c0:c1] = dn
            prev = (d, hit, xf, Rf)
    return SyntheticSequence(depth=depth, flow=flow, mask=mask, pose=pose, pose_valid=pose_valid,
                             gt_pose=gt_pose, gt_twist=gt_twist, flow_grid=grid, flow_scale=scale, dt=dt, half=half_used)
