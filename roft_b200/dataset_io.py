"""Fast-YCB-format sequence directories: writers/readers used by the host-side tests and tools.

    <seq>/depth/<i>.float              u64 width, u64 height, H*W float32 metres (tools/dataset/conversion/ho3d_utils.py:74-79)
    <seq>/optical_flow/<set>/<i>.float i32 cv_type (13 CV_32FC2 | 11 CV_16SC2), u64 cols, u64 rows, data
                                       (src/roft-lib/src/OpticalFlowUtilities.cpp:38-62,99-119); no file for frame 0
    <seq>/masks/<set>/<object>_<i>.pgm 8-bit mask (the reference reads .png; PNG coding is host IO outside the scope)
    <seq>/<set>/poses.txt              "x y z ax ay az angle" per frame, all-zero row = invalid
    <seq>/data.txt                     "stamp_rgb stamp_depth x y z ax ay az angle" (camera pose)
    <seq>/cam_K.json                   width, height, fx, fy, cx, cy
"""
from __future__ import annotations

import json
import os
import struct

import numpy as np

CV_32FC2 = 13
CV_16SC2 = 11


def write_depth(path: str, depth: np.ndarray) -> None:
    h, w = depth.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<QQ", w, h))
        f.write(np.ascontiguousarray(depth, np.float32).tobytes())


def read_depth(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        w, h = struct.unpack("<QQ", f.read(16))
        return np.frombuffer(f.read(), np.float32).reshape(h, w)


def write_flow(path: str, flow: np.ndarray) -> None:
    rows, cols, _ = flow.shape
    cv_type = CV_16SC2 if flow.dtype == np.int16 else CV_32FC2
    with open(path, "wb") as f:
        f.write(struct.pack("<i", cv_type))
        f.write(struct.pack("<QQ", cols, rows))
        f.write(np.ascontiguousarray(flow).tobytes())


def read_flow(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        (cv_type,) = struct.unpack("<i", f.read(4))
        cols, rows = struct.unpack("<QQ", f.read(16))
        dt = np.int16 if cv_type == CV_16SC2 else np.float32
        return np.frombuffer(f.read(), dt).reshape(rows, cols, 2)


def write_pgm(path: str, mask: np.ndarray) -> None:
    h, w = mask.shape
    with open(path, "wb") as f:
        f.write(f"P5\n{w} {h}\n255\n".encode())
        f.write(np.ascontiguousarray(mask, np.uint8).tobytes())


def quat_to_axis_angle(q: np.ndarray) -> np.ndarray:
    """(w,x,y,z) -> (ax, ay, az, angle) like Eigen::AngleAxisd(Quaterniond)."""
    n = float(np.linalg.norm(q[1:]))
    if n == 0:
        return np.array([1.0, 0.0, 0.0, 0.0])
    angle = 2.0 * np.arctan2(n, abs(q[0]))
    if q[0] < 0:
        n = -n
    return np.array([q[1] / n, q[2] / n, q[3] / n, angle])


def write_sequence(root: str, seq, track: int, object_name: str = "003_cracker_box", flow_set: str = "nvof",
                   mask_set: str = "gt", pose_set: str = "gt", fx=None, fy=None, cx=None, cy=None, mask_format: str = "pgm") -> None:
    """Dump one track of a roft_b200.synthetic.SyntheticSequence as a Fast-YCB-format directory."""
    F = seq.depth.shape[0]
    H, W = seq.depth.shape[2], seq.depth.shape[3]
    for d in ("depth", f"optical_flow/{flow_set}", f"masks/{mask_set}", pose_set):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    with open(os.path.join(root, "data.txt"), "w") as fd, open(os.path.join(root, pose_set, "poses.txt"), "w") as fp:
        for k in range(F):
            write_depth(os.path.join(root, "depth", f"{k}.float"), seq.depth[k, track].cpu().numpy())
            if k > 0:
                write_flow(os.path.join(root, "optical_flow", flow_set, f"{k}.float"), seq.flow[k, track].cpu().numpy())
            if mask_format == "png":  # the format of the Fast-YCB / HO-3D mask directories (needs OpenCV for the writer only)
                import cv2
                cv2.imwrite(os.path.join(root, "masks", mask_set, f"{object_name}_{k}.png"), seq.mask[k, track].cpu().numpy())
            else:
                write_pgm(os.path.join(root, "masks", mask_set, f"{object_name}_{k}.pgm"), seq.mask[k, track].cpu().numpy())
            stamp = k * seq.dt
            fd.write(f"{stamp!r} {stamp!r} 0 0 0 1 0 0 0\n")
            p = seq.pose[k, track].numpy()
            if bool(seq.pose_valid[k, track]):
                aa = quat_to_axis_angle(p[3:])
                fp.write(" ".join(repr(float(v)) for v in list(p[:3]) + list(aa)) + "\n")
            else:
                fp.write("0 0 0 0 0 0 0\n")
    with open(os.path.join(root, "cam_K.json"), "w") as f:
        json.dump({"name": "synthetic", "width": W, "height": H, "fx": fx, "fy": fy, "cx": cx, "cy": cy}, f, indent=1)


def write_obj(path: str, vertices, faces) -> None:
    """Wavefront OBJ (v / f records, 1-based indices): the mesh input of the render-and-compare pose test."""
    with open(path, "w") as f:
        for v in np.asarray(vertices, np.float64):
            f.write(f"v {float(v[0])!r} {float(v[1])!r} {float(v[2])!r}\n")
        for t in np.asarray(faces, np.int64):
            f.write(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n")
