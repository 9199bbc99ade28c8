#!/usr/bin/env python
"""Benchmark of the ROFT hot path (BASELINE.json: tracked frames/s at 1280x720, batched tracks).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

One "step" = ROFTFilter::filtering_step for every track of the batch: flow->velocity measurement with
Laplacian re-weighting and the velocity Kalman correction, flow-aided mask synchronisation (delayed
masks), and the pose UKF with delayed pose measurements and re-synchronisation.

Workloads (--workload; all 1280x720, 256 tracks per GPU, synthetic Fast-YCB-format data):
  c4   BASELINE configs[3] (default): dense CV_32FC2 flow, every masked pixel (subsampling_radius 1), mask
       coverage ~0.25, masks / poses delayed by 6 frames - the full configs[1] pipeline per track
  c5   BASELINE configs[4]: large masks (coverage >= 0.40), 4-frame mask-sync delay; 2048 tracks = 256 per GPU
       at 8 GPUs; with fewer GPUs the per-GPU batch stays 256 (weak scaling)
  ref  the reference's own default configuration (test/test.sh:63, config_fast_ycb.cfg:83): CV_16SC2 flow on
       a 4-pixel grid (NVOF 1.0, S10.5), subsampling_radius 35
  full coverage 1.0 (every pixel of the frame masked): the worst case for the worklist

`value`   : tracked frames/s with all inputs already resident in HBM (zero-copy device frames).
`e2e`     : the same metric through the C ABI with HOST (pinned) buffers: every step copies that step's
            depth/flow(/mask) host->device and reads the beliefs back device->host.
`roofline`: algorithmic bytes of a step (W*H*(4 + flow + 1) per track-frame) over its
            CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth.
`cpu_baseline`: the reference algorithm (oracle/cpu_ref.cpp, sequential SKF as in SKFCorrection.cpp) on the
            host cores, on a bounded sample of the same workload.
`sanity.parity`: a few randomly chosen tracks of the batch replayed through the CPU restatement outside the
            timed region (mask bit-equal, velocity / pose relative difference).
`extras`  (N = 1): the other workloads (short runs) and the configs[2] latency block (640x480, one track, host
            buffers, p50 / p95 over 300 frames).
Multi-GPU: independent tracks are partitioned across ranks (no data-path collective); weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1280, 720

WORKLOADS = {
    #        coverage, delay, stride, flow format
    "c4":   dict(coverage=0.25, delay=6, stride=1, fmt="f32",
                 name="BASELINE configs[3]: dense CV_32FC2 flow, mask coverage ~0.25, subsampling_radius 1, mask+pose delay 6"),
    "c5":   dict(coverage=0.40, delay=4, stride=1, fmt="f32",
                 name="BASELINE configs[4]: dense CV_32FC2 flow, large masks (coverage >= 0.40), subsampling_radius 1, mask+pose delay 4"),
    "ref":  dict(coverage=0.25, delay=6, stride=35, fmt="s16",
                 name="reference defaults: CV_16SC2 flow on a 4-px grid (NVOF 1.0), subsampling_radius 35, mask coverage ~0.25, delay 6"),
    "full": dict(coverage=1.0, delay=6, stride=1, fmt="f32",
                 name="every pixel masked (coverage 1.0), dense CV_32FC2 flow, subsampling_radius 1, delay 6"),
    "c2":   dict(coverage=0.25, delay=6, stride=1, fmt="f32", outlier_rejection=True,
                 name="BASELINE configs[1] per track: c4 + render-and-compare pose outlier rejection on every re-synchronisation "
                      "(cuboid mesh, per-track scale)"),
}


def bytes_per_track_frame(fmt):
    """depth f32 + flow + mask u8, each input byte counted once."""
    flow = W * H * 8 if fmt == "f32" else (W // 4) * (H // 4) * 4
    return W * H * 4 + flow + W * H


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="roft_b200", choices=["roft_b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--tracks", type=int, default=256, help="tracks per GPU")
    ap.add_argument("--frames", type=int, default=12, help="distinct resident frames per track")
    ap.add_argument("--stride", type=int, default=None, help="override the workload's subsampling radius")
    ap.add_argument("--coverage", type=float, default=None, help="override the workload's target mask coverage")
    ap.add_argument("--delay", type=int, default=None, help="override the workload's mask / pose delay in frames")
    ap.add_argument("--accum", default="auto", choices=["auto", "fp64", "fp32"],
                    help="precision of the per-pixel terms of the normal equations (roftb_config.accum_fp64)")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sweep", "--no-extras", dest="no_extras", action="store_true",
                    help="skip the extra workloads and the latency block (N = 1 only)")
    ap.add_argument("--no-parity", action="store_true", help="skip the CPU replay of sample tracks")
    ap.add_argument("--parity-tracks", type=int, default=4)
    ap.add_argument("--parity-steps", type=int, default=14)
    ap.add_argument("--no-resync", action="store_true", help="diagnostic: pose re-sync replay off")
    ap.add_argument("--per-step", action="store_true", help="diagnostic: print main-stream ms per step by phase of the mask period")
    ap.add_argument("--single-mask", action="store_true", help="diagnostic: deliver the mask / pose only at step 0")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    a.fmt = wl["fmt"]
    a.stride = wl["stride"] if a.stride is None else a.stride
    a.coverage = wl["coverage"] if a.coverage is None else a.coverage
    a.delay = wl["delay"] if a.delay is None else a.delay
    return a


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 2 ms from before the warm-up (the
    timed region of a default run is a few tens of ms - shorter than nvidia-smi's start-up, so the CLI loop is only
    the fallback, and a one-shot query right after the region the last resort)."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []   # (t, sm_mhz, max_mhz, set(reasons))
        self.stop_flag = threading.Event()
        self.proc = None
        self.errors = []

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        # NVML indexes the devices the container exposes; CUDA_VISIBLE_DEVICES may renumber them for CUDA
        count = nv.nvmlDeviceGetCount()
        cand = []
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                cand.append(int(ids[self.index]))
        cand += [self.index, 0]
        h = None
        for c in cand:
            if 0 <= c < count:
                try:
                    h = nv.nvmlDeviceGetHandleByIndex(c)
                    break
                except Exception:
                    continue
        if h is None:
            raise RuntimeError("no NVML device")
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.samples.append((time.perf_counter(), float(sm), float(mx), {n for n, b in bits.items() if r & b}))
            time.sleep(0.002)

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def _parse_smi(self, line):
        s = [x.strip() for x in line.split(",")]
        return (time.perf_counter(), float(s[0]), float(s[1]),
                {n for n, v in zip(self.NAMES, s[3:7]) if v.lower().startswith("active")})

    def _run_smi(self):
        self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            if self.stop_flag.is_set():
                break
            try:
                self.samples.append(self._parse_smi(line))
            except Exception:
                continue

    def one_shot(self):
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
            if out:
                self.samples.append(self._parse_smi(out[0]))
        except Exception as e:
            self.errors.append("one-shot nvidia-smi: " + repr(e))

    def run(self):
        try:
            self._run_nvml()
        except Exception as e:
            self.errors.append("nvml: " + repr(e))
            try:
                self._run_smi()
            except Exception as e2:
                self.errors.append("nvidia-smi loop: " + repr(e2))

    def finish(self, t0=None, t1=None):
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        inside = [s for s in self.samples if t0 is None or (t0 <= s[0] <= t1)]
        where = "timed region"
        if not inside and self.samples:  # region shorter than the sampling period: the samples closest to it
            mid = 0.5 * (t0 + t1)
            inside = sorted(self.samples, key=lambda s: abs(s[0] - mid))[:4]
            where = "closest to the timed region"
        if not inside:
            self.one_shot()
            inside = self.samples[-1:]
            where = "one-shot query right after the timed region"
        sm = sorted(s[1] for s in inside)
        reasons = set()
        for s in inside:
            reasons |= s[3]
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((s[2] for s in inside), default=None),
               "reasons": sorted(reasons), "samples": len(sm), "sampled": where}
        if self.errors:
            out["sampler_notes"] = self.errors
        return out


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(key: str, tracks: int):
    """DRAM bytes per launch from the committed `ncu --set full` capture (profiles/roofline_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            d = json.load(f)
        e = d.get(key)
        if e and e.get("tracks") == tracks:
            return e["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def workload_name(args):
    return (f"{args.tracks} independent tracks/GPU, 1280x720, {WORKLOADS[args.workload]['name']}; Laplacian weighting on, "
            f"flow-aided mask sync and pose re-sync (workload '{args.workload}': coverage {args.coverage:.2f}, "
            f"subsampling_radius {args.stride}, delay {args.delay}, flow {args.fmt})")


class Runner:
    """One tracker + its resident synthetic frames for a workload."""

    def __init__(self, api, dev, local, T, F, coverage, delay, stride, fmt, accum, first_track=0, no_resync=False,
                 single_mask=False, width=W, height=H, intr=None, outlier_rejection=False):
        import numpy as np
        import torch
        from roft_b200.synthetic import make_sequence
        self.np, self.torch = np, torch
        self.T, self.F, self.D, self.single_mask = T, F, delay, single_mask
        kw = {}
        if intr:
            kw.update(fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3])
        cfg = api.default_config(n_tracks=T, width=width, height=height, subsampling_radius=stride, segm_delay=delay,
                                 pose_delay=delay, device=local, accum_fp64={"fp32": 0, "fp64": 1, "auto": 2}[accum],
                                 flow_format=api.FLOW_F32 if fmt == "f32" else api.FLOW_S16,
                                 flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0,
                                 **({"use_pose_resync": 0} if no_resync else {}),
                                 **({"outlier_rejection": 1} if outlier_rejection else {}), **kw)
        self.cfg = cfg
        self.trk = api.Tracker(cfg)
        # frames 0..F; frame 0 only provides the initial mask/pose, frames 1..F are cycled
        skw = dict(fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3]) if intr else {}
        self.seq = make_sequence(T, F + 1, width, height, device=dev, target_coverage=coverage, first_track_id=first_track,
                                 track_chunk=8, flow_format=fmt, **skw)
        torch.cuda.synchronize()
        self.x0 = np.zeros((T, 13)); self.x0[:, 6:] = self.seq.pose[0].numpy()
        self.pose_np = self.seq.pose.numpy(); self.pv_np = self.seq.pose_valid.numpy().astype(np.uint8)
        if outlier_rejection:  # every synthetic track shows a cuboid: unit mesh + the track's half extents
            from roft_b200.synthetic import cuboid_mesh
            self.trk.set_mesh(*cuboid_mesh([1.0, 1.0, 1.0]))
            self.trk.set_mesh_scale(self.seq.half.numpy().astype(np.float32))
        self.trk.init(self.x0)
        self.step_i = 0

    def frame_of(self, step):  # 0, then 1..F cycled
        return 0 if step == 0 else 1 + (step - 1) % self.F

    def stale(self, step):  # DatasetImageSegmentationDelayed.cpp:42-63: frame delivered (late) at this step, or None
        idx = step - self.D
        if idx % self.D != 0 or (self.single_mask and step > 0):
            return None
        return self.frame_of(max(idx, 0))

    def do_step(self, host=None):
        step = self.step_i
        f = self.frame_of(step)
        s = self.stale(step)
        pose = self.pose_np[s] if s is not None else None
        pv = self.pv_np[s] if s is not None else None
        seq = self.seq
        if host is None:
            self.trk.step(seq.depth[f], seq.flow[f] if step > 0 else None, seq.mask[s] if s is not None else None,
                          pose=pose, pose_valid=pv, device=True)
        else:
            hd, hf, hm = host
            self.trk.step(hd[f % len(hd)], hf[f % len(hf)] if step > 0 else None, hm[0] if s is not None else None,
                          pose=pose, pose_valid=pv, device=False)
        self.step_i += 1

    def reset(self):
        self.trk.init(self.x0)
        self.step_i = 0


def timed_resident(r: Runner, steps, warmup, barrier, dev, per_step=False):
    """W warm-up steps, then K steps between CUDA events on the library's main stream (joined with its other streams)."""
    torch = r.torch
    trk = r.trk
    for _ in range(warmup):
        r.do_step()
    barrier()
    trk.profile(True)
    l0 = trk.kernel_launches
    barrier()
    t0 = time.perf_counter()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ext = torch.cuda.ExternalStream(trk.stream, device=dev)
    ev0.record(ext)
    step_ev = []
    host_t = [time.perf_counter()]
    for _ in range(steps):
        r.do_step()
        host_t.append(time.perf_counter())
        if per_step:  # diagnostic: main-stream time stamps per step (velocity chain only)
            e = torch.cuda.Event(enable_timing=True); e.record(ext); step_ev.append((r.step_i - 1, e))
    host_issue_ms = (host_t[-1] - host_t[0]) * 1e3 / steps
    trk.join()
    ev1.record(ext)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    launches = trk.kernel_launches - l0
    phases, _ = trk.profile(False)
    if per_step:
        D = r.D
        prev = ev0
        per = {}
        for st_i, e in step_ev:
            per.setdefault((st_i - D) % D, []).append(prev.elapsed_time(e)); prev = e
        print("per-step ms by (step-D)%D:", {k: round(sum(v) / len(v), 3) for k, v in sorted(per.items())}, file=sys.stderr)
        hper = {}
        for i, (st_i, _) in enumerate(step_ev):
            hper.setdefault((st_i - D) % D, []).append((host_t[i + 1] - host_t[i]) * 1e3)
        print("host issue ms by (step-D)%D:", {k: round(sum(v) / len(v), 3) for k, v in sorted(hper.items())}, file=sys.stderr)
    return dict(dev_ms=dev_ms, wall=wall, t0=t0, launches=launches, phases=phases, host_issue_ms=host_issue_ms)


def parity_sample(args, api, dev, local, r: Runner, rank):
    """Replay a few randomly chosen tracks of the batch through the CPU restatement (oracle/cpu_ref) for the first
    steps of the run - outside every timed region - and compare mask (bit-equal), velocity and pose each step."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_ref  # noqa: checker only (tests / smoke / this sanity block / the cpu_baseline leg)
    import roft_oracle as o
    rng = np.random.default_rng(1234 + rank)
    K = min(args.parity_tracks, r.T)
    picks = sorted(rng.choice(r.T, size=K, replace=False).tolist())
    fmt = args.fmt
    ocfg = o.RoftConfig(width=W, height=H, subsampling_radius=float(args.stride), segm_delay=args.delay, pose_delay=args.delay,
                        flow_grid=1 if fmt == "f32" else 4, flow_scale=1.0 if fmt == "f32" else 32.0,
                        use_pose_resync=not args.no_resync)
    filters = [cpu_ref.CFilter(ocfg, r.x0[t]) for t in picks]
    r.reset()
    seq = r.seq
    worst = {"velocity": 0.0, "position": 0.0, "quaternion": 0.0}
    mask_equal = True
    counts_equal = True

    def rel(a, b):
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-9))

    for step in range(args.parity_steps):
        f = r.frame_of(step)
        s = r.stale(step)
        r.do_step()
        pm, vm = r.trk.state()
        cnt, _, _ = r.trk.velocity_info()
        raw, _ = r.trk.mask(raw=True, thresholded=False)
        for k, t in enumerate(picks):
            pose = r.pose_np[s, t] if (s is not None and r.pv_np[s, t]) else None
            filters[k].step(seq.depth[f, t].cpu().numpy(), seq.flow[f, t].cpu().numpy() if step > 0 else None,
                            seq.mask[s, t].cpu().numpy() if s is not None else None, pose)
            cpm, _, cvm, _, cn = filters[k].state()
            mask_equal &= bool(np.array_equal(filters[k].mask()[0], raw[t]))
            counts_equal &= (int(cnt[t]) == int(cn)) if step > 0 else True
            if np.linalg.norm(cvm) > 1e-9:
                worst["velocity"] = max(worst["velocity"], rel(vm[t], cvm))
            worst["position"] = max(worst["position"], rel(pm[t, 6:9], cpm[6:9]))
            worst["quaternion"] = max(worst["quaternion"], float(min(np.linalg.norm(pm[t, 9:] - cpm[9:]), np.linalg.norm(pm[t, 9:] + cpm[9:]))))
    r.reset()
    return {"tracks": picks, "steps": args.parity_steps, "mask_bit_equal": mask_equal, "valid_pixel_counts_equal": counts_equal,
            "max_rel": worst, "parity_max_rel": max(worst.values()), "tolerance": 1e-4,
            "checker": "oracle/cpu_ref (sequential per-pixel SKF, FP64)"}


def latency_block(api, dev, local, frames=300):
    """BASELINE configs[2]: HO-3D-format 640x480, ONE track, every masked pixel, host buffers through the C ABI:
    per frame = upload + step + read-back of the beliefs (blocking).  p50 / p95 over `frames` frames."""
    import numpy as np
    r = Runner(api, dev, local, 1, 12, 0.25, 6, 1, "f32", "auto", width=640, height=480, intr=(617.0, 617.0, 312.0, 241.0))
    seq = r.seq
    # pinned host buffers, like the e2e measurement (a caller that hands over pageable memory pays the driver's staging copy)
    hd = [seq.depth[i].cpu().pin_memory() for i in range(r.F + 1)]
    hf = [seq.flow[i].cpu().pin_memory() for i in range(r.F + 1)]
    hm = [seq.mask[i].cpu().pin_memory() for i in range(r.F + 1)]
    lat = []
    for step in range(frames + 20):
        f = r.frame_of(step); s = r.stale(step)
        t0 = time.perf_counter()
        r.trk.step(hd[f], hf[f] if step > 0 else None, hm[s] if s is not None else None,
                   pose=r.pose_np[s] if s is not None else None, pose_valid=r.pv_np[s] if s is not None else None, device=False)
        pm, vm = r.trk.state()
        if step >= 20:
            lat.append((time.perf_counter() - t0) * 1e3)
        r.step_i += 1
    lat = np.sort(np.array(lat))
    return {"workload": "BASELINE configs[2]: 640x480, single track, dense flow, subsampling_radius 1, delay 6, pinned host buffers "
                        "(upload + step + blocking read-back of the beliefs per frame)",
            "frames": int(len(lat)), "p50_ms": float(lat[len(lat) // 2]), "p95_ms": float(lat[int(len(lat) * 0.95)]),
            "mean_ms": float(lat.mean()), "finite": bool(np.isfinite(pm).all() and np.isfinite(vm).all())}


def extra_workload(api, dev, local, name, T, peak, accum):
    """Short device-resident run of another workload (reported beside the headline, never instead of it)."""
    import numpy as np
    import torch
    wl = WORKLOADS[name.split("@")[0]]  # "c4@512": the c4 workload at 512 tracks per GPU
    r = Runner(api, dev, local, T, 6, wl["coverage"], wl["delay"], wl["stride"], wl["fmt"], accum,
               outlier_rejection=wl.get("outlier_rejection", False))

    def barrier():
        torch.cuda.synchronize(); r.trk.sync()

    m = timed_resident(r, 24, 12, barrier, dev)
    ms = m["dev_ms"] / 24
    B = bytes_per_track_frame(wl["fmt"])
    gbs = T * B / (ms * 1e-3) / 1e9
    units, pixels = r.trk.worklist()
    pm, vm = r.trk.state()
    cnt, _, _ = r.trk.velocity_info()
    out = {"workload": name, "description": wl["name"], "tracks_per_gpu": T, "steps": 24, "warmup": 12,
           "value": T / (ms * 1e-3), "unit": "tracked frames/s", "ms_per_step": ms,
           "algorithmic_bytes_per_track_frame": B, "roofline_achieved": gbs, "roofline_frac": gbs / peak,
           "mean_listed_units": float(units.mean()), "mean_valid_pixels": float(cnt.mean()),
           # dram__bytes_read + dram__bytes_write of ONE velocity-kernel launch of this workload (committed ncu capture), or null
           "velocity_kernel_traffic": ncu_traffic(f"velocity_{name.split('@')[0]}", T),
           "finite": bool(np.isfinite(pm).all() and np.isfinite(vm).all())}
    del r
    torch.cuda.empty_cache()
    return out


def run_own(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from roft_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (roft_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    # one rank per GPU: keep this rank's host thread on its own cores so eight launch loops do not migrate over each other
    try:
        ncpu = os.cpu_count() or 1
        if world > 1 and hasattr(os, "sched_setaffinity") and ncpu >= 2 * world:
            per = ncpu // world
            os.sched_setaffinity(0, set(range(local * per, (local + 1) * per)))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"
    T = args.tracks
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    r = Runner(api, dev, local, T, args.frames, args.coverage, args.delay, args.stride, args.fmt, args.accum,
               outlier_rejection=WORKLOADS[args.workload].get("outlier_rejection", False),
               first_track=rank * T, no_resync=args.no_resync, single_mask=args.single_mask)
    trk = r.trk

    def barrier():
        torch.cuda.synchronize()
        trk.sync()
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput ------------------------------------------------------
    m = timed_resident(r, args.steps, args.warmup, barrier, dev, per_step=args.per_step and rank == 0)
    clocks = sampler.finish(m["t0"], m["t0"] + m["wall"]) if sampler else None
    mine = torch.tensor([m["dev_ms"], m["host_issue_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        per_rank = [[float(v[0]), float(v[1])] for v in allv]
    else:
        per_rank = [[float(mine[0]), float(mine[1])]]
    ms_total = max(v[0] for v in per_rank)
    ms_per_step = ms_total / args.steps
    value = world * T * args.steps / (ms_total * 1e-3)

    # sanity: the tracker must actually track (guards against a silently skipped pipeline)
    pm, vm = trk.state()
    cnt, _, _ = trk.velocity_info()
    units, pixels = trk.worklist()

    # ---- end to end through the C ABI with host buffers -----------------------------------
    e2e = None
    if not args.no_e2e:
        nh = 2
        Hf, Wf = (H, W) if args.fmt == "f32" else (H // 4, W // 4)
        fdt = torch.float32 if args.fmt == "f32" else torch.int16
        hd = [torch.empty((T, H, W), dtype=torch.float32).pin_memory() for _ in range(nh)]
        hf = [torch.empty((T, Hf, Wf, 2), dtype=fdt).pin_memory() for _ in range(nh)]
        hm = [torch.empty((T, H, W), dtype=torch.uint8).pin_memory()]
        for i in range(nh):
            hd[i].copy_(r.seq.depth[1 + i]); hf[i].copy_(r.seq.flow[1 + i])
        hm[0].copy_(r.seq.mask[1])
        host = ([x.numpy() for x in hd], [x.numpy() for x in hf], [x.numpy() for x in hm])
        r.reset()
        r.do_step(host); r.do_step(host)
        trk.state()
        barrier()
        t0 = time.perf_counter()
        h2d = 0
        fbytes = int(hf[0][0].numel()) * hf[0].element_size()
        for _ in range(args.e2e_steps):
            has_mask = r.stale(r.step_i) is not None
            r.do_step(host)
            h2d += T * (H * W * 4 + fbytes) + (T * H * W if has_mask else 0)
            trk.state()  # device->host read of the step's result (pose 13 + velocity 6 doubles per track)
        barrier()
        e2e_s = time.perf_counter() - t0
        ts = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e = {"value": world * T * args.e2e_steps / float(ts.item()), "unit": "tracked frames/s",
               "h2d_bytes_per_step": world * h2d // args.e2e_steps, "d2h_bytes_per_step": world * T * 19 * 8,
               "steps": args.e2e_steps, "h2d_gbs_per_rank": h2d / float(ts.item()) / 1e9}

    # ---- parity of sample tracks against the CPU restatement (outside the timed regions) ------------------
    parity = None
    if WORKLOADS[args.workload].get("outlier_rejection"):
        # the C++ checker has no rasteriser: this workload's parity is the GPU tests' (numpy restatement, small frames)
        parity = {"skipped": "outlier rejection is checked by tests/test_gpu_parity.py against oracle/roft_oracle.py"}
    elif not args.no_parity and rank == 0:
        try:
            parity = parity_sample(args, api, dev, local, r, rank)
        except Exception as e:  # a checker problem must not lose the bench line; it is reported
            parity = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_kind = measured_peak_gbs()
    B = bytes_per_track_frame(args.fmt)
    # Roofline.  The unit of SURVEY 8(d) is a track-frame (depth + flow + mask, each input byte counted once); it is
    # consumed by one launch of the velocity kernel (plus the small kernels around it), so `achieved` = algorithmic
    # bytes of the step / device time of the step.  The velocity kernel itself is reported beside it with the bytes
    # its pass A has to touch: the worklist restricts it to the non-empty 128-px units of the mask.
    phases = m["phases"]
    vel_ms = phases["flow_pass_a"] + phases["median_select"] + phases["flow_pass_b"] + phases["epilogue"]
    unit_bytes = 128 * (1 + 4 + 8) if args.fmt == "f32" else None
    dom_bytes = int(units.astype(np.int64).sum()) * unit_bytes if unit_bytes else None
    dom_gbs = dom_bytes / (vel_ms * 1e-3) / 1e9 if (dom_bytes and vel_ms > 0) else None
    achieved = T * B / (ms_per_step * 1e-3) / 1e9
    out = {
        "metric": "tracked frames/sec at 1280x720 (batched tracks)", "value": value, "unit": "tracked frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"auto": "f32 per-pixel terms + f64 reduction/solve (f64 terms for tracks < 32768 px)", "fp64": "f64", "fp32": "f32"}[args.accum],
        "data": "synthetic",
        "config": {"workload": workload_name(args), "workload_key": args.workload, "tracks_per_gpu": T, "resident_frames": args.frames,
                   "l2_policy": f"inputs larger than L2: {T * B / 1e9:.2f} GB of frame data per step, no flush needed",
                   "accumulation": args.accum, "parallelism": f"tracks partitioned over {world} GPU(s), no collective",
                   "note_c5": "2048 tracks = 256 per GPU at 8 GPUs; the per-GPU batch stays 256 for N < 8 (weak scaling)"
                   if args.workload == "c5" else None},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("step_" + args.workload, T), "peak_kind": peak_kind,
                     "scope": "whole step: algorithmic bytes of T track-frames / device time of the step (all kernels, all streams)",
                     "algorithmic_bytes_per_track_frame": B, "algorithmic_bytes_per_step": T * B,
                     "dominant_kernel": {"name": "k_velocity_track", "ms": vel_ms, "listed_units_mean": float(units.mean()),
                                         "algorithmic_bytes": dom_bytes, "achieved": dom_gbs,
                                         "frac": (dom_gbs / peak) if dom_gbs else None,
                                         "traffic": ncu_traffic("velocity_" + args.workload, T),
                                         "note": "bytes = listed 128-px units x (mask + depth + flow); duration from CUDA events "
                                                 "inside the overlapped step; profiles/ holds the isolated ncu duration"}},
        "phases_ms_per_step": phases,
        "per_rank": {"ms_per_step": [round(v[0] / args.steps, 5) for v in per_rank],
                     "host_issue_ms_per_step": [round(v[1], 4) for v in per_rank]},
        "gpu_launches": int(m["launches"]),
        "host_issue_ms_per_step": round(m["host_issue_ms"], 4),
        "clocks": clocks,
        "e2e": e2e,
        "wall_s": m["wall"],
        "sanity": {"mean_valid_pixels": float(cnt.mean()), "mean_listed_units": float(units.mean()),
                   "mean_abs_w": float(np.abs(vm[:, 3:]).mean()),
                   "finite": bool(np.isfinite(pm).all() and np.isfinite(vm).all()),
                   "parity": parity, "parity_max_rel": (parity or {}).get("parity_max_rel")},
    }
    if not args.no_cpu and world == 1:  # reported at N = 1 only (rank 0)
        try:
            out["cpu_baseline"] = cpu_baseline(args, r.seq, "port")
        except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
            out["cpu_baseline"] = {"error": repr(e)}
    if world == 1 and not args.no_extras and not args.single_mask and args.workload == "c4":
        extras = {}
        del r, trk
        torch.cuda.empty_cache()
        for name in ("c2", "c5", "full", "ref", "c4@512"):
            try:
                extras[name] = extra_workload(api, dev, local, name, int(name.split("@")[1]) if "@" in name else T, peak, args.accum)
            except Exception as e:
                extras[name] = {"error": repr(e)}
        try:
            extras["latency_c3"] = latency_block(api, dev, local)
        except Exception as e:
            extras["latency_c3"] = {"error": repr(e)}
        out["extras"] = extras
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, seq, kind, threads=None, seconds=None):
    """Time the CPU restatement of the reference (oracle/cpu_ref) on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_ref  # noqa: test infrastructure, timed here as the reported baseline only
    return cpu_ref.timed_baseline(seq, stride=args.stride, delay=args.delay, seconds=seconds or args.cpu_seconds,
                                  threads=threads, kind=kind)


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on the host cores (all threads), same config."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from roft_b200.synthetic import make_sequence
    ncpu = os.cpu_count() or 1
    n_tracks = max(1, min(ncpu, 32))
    seq = make_sequence(n_tracks, min(args.frames, 6) + 1, W, H, device="cpu", target_coverage=args.coverage, track_chunk=4,
                        flow_format=args.fmt)
    t0 = time.perf_counter()
    cb = cpu_baseline(args, seq, "port", threads=ncpu, seconds=max(10.0, args.cpu_seconds))
    out = {
        "impl": "reference", "metric": "tracked frames/sec at 1280x720 (batched tracks)", "value": cb["value"],
        "unit": "tracked frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * n_tracks / cb["value"] if cb["value"] else None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "workload_key": args.workload,
                   "note": "CPU restatement of the reference (the reference itself needs Eigen/OpenCV/bfl, absent here); "
                           "one thread per track over all host cores; a step is a bounded sample of the workload"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "tracked frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_own(a)
