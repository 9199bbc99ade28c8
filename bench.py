#!/usr/bin/env python
"""Benchmark of the ROFT hot path (BASELINE.json: tracked frames/s at 1280x720, batched tracks).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

One "step" = ROFTFilter::filtering_step for every track of the batch: flow->velocity measurement
with Laplacian re-weighting and the velocity Kalman correction, flow-aided mask synchronisation
(delayed masks), and the pose UKF with delayed pose measurements and re-synchronisation.

`value`   : tracked frames/s with all inputs already resident in HBM (zero-copy device frames).
`e2e`     : the same metric through the C ABI with HOST (pinned) buffers: every step copies that
            step's depth/flow(/mask) host->device and reads the beliefs back device->host.
`roofline`: dominant kernel's algorithmic bytes (W*H*(4+8+1) per track-frame) over its
            CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth.
`cpu_baseline`: the reference algorithm (oracle/cpu_ref.cpp, sequential SKF as in SKFCorrection.cpp)
            on the host cores, on a bounded sample of the same workload.
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
BYTES_PER_TRACK_FRAME = W * H * (4 + 8 + 1)  # depth f32 + dense flow 2xf32 + mask u8, each counted once


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="roft_b200", choices=["roft_b200", "reference"])
    ap.add_argument("--tracks", type=int, default=256, help="tracks per GPU")
    ap.add_argument("--frames", type=int, default=12, help="distinct resident frames per track")
    ap.add_argument("--stride", type=int, default=1, help="subsampling radius (reference default 35; 1 = every masked pixel)")
    ap.add_argument("--coverage", type=float, default=0.25, help="target mask coverage of the frame")
    ap.add_argument("--delay", type=int, default=6, help="mask / pose delay in frames")
    ap.add_argument("--accum", default="auto", choices=["auto", "fp64", "fp32"],
                    help="precision of the per-pixel terms of the normal equations (roftb_config.accum_fp64)")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the extra 512-tracks-per-GPU measurement (N=1 only)")
    ap.add_argument("--no-resync", action="store_true", help="diagnostic: pose re-sync replay off")
    ap.add_argument("--per-step", action="store_true", help="diagnostic: print main-stream ms per step by phase of the mask period")
    ap.add_argument("--single-mask", action="store_true", help="diagnostic: deliver the mask / pose only at step 0")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (NVML every 2 ms; the timed region of a default
    run is ~0.15 s, shorter than nvidia-smi's start-up, so the CLI loop is only the fallback)."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []   # (t, sm_mhz, max_mhz, set(reasons))
        self.stop_flag = threading.Event()
        self.proc = None

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

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            if self.stop_flag.is_set():
                break
            s = [x.strip() for x in line.split(",")]
            try:
                self.samples.append((time.perf_counter(), float(s[0]), float(s[1]),
                                     {n for n, v in zip(self.NAMES, s[3:7]) if v.lower().startswith("active")}))
            except Exception:
                continue

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                pass

    def finish(self, t0=None, t1=None):
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        inside = [s for s in self.samples if t0 is None or (t0 <= s[0] <= t1)]
        where = "timed region"
        if not inside and self.samples:  # region shorter than the sampling period: take the closest samples
            inside = self.samples[-2:]
            where = "closest to the timed region"
        sm = sorted(s[1] for s in inside)
        reasons = set()
        for s in inside:
            reasons |= s[3]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((s[2] for s in inside), default=None),
                "reasons": sorted(reasons), "samples": len(sm), "sampled": where}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel: str, tracks: int, fp32: bool):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            d = json.load(f)
        e = d.get(kernel + ("_fp32" if fp32 and kernel in ("flow_pass_b", "step") else ""))
        if e and e.get("tracks") == tracks:
            return e["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def workload_name(args):
    return (f"{args.tracks} independent tracks/GPU, 1280x720, dense CV_32FC2 flow, mask coverage ~{args.coverage:.2f}, "
            f"subsampling_radius {args.stride}, Laplacian weighting on, mask+pose delay {args.delay} frames with "
            f"flow-aided sync and pose re-sync (BASELINE configs[3], full configs[1] pipeline per track)")


def build_frames(args, device, first_track):
    import torch
    from roft_b200.synthetic import make_sequence
    # frames 0..F; frame 0 only provides the initial mask/pose, frames 1..F are cycled
    seq = make_sequence(args.tracks, args.frames + 1, W, H, device=device, target_coverage=args.coverage,
                        first_track_id=first_track, track_chunk=8)
    torch.cuda.synchronize()
    return seq


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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"
    T, F, D = args.tracks, args.frames, args.delay

    cfg = api.default_config(n_tracks=T, subsampling_radius=args.stride, segm_delay=D, pose_delay=D, device=local,
                             accum_fp64={"fp32": 0, "fp64": 1, "auto": 2}[args.accum],
                             **({"use_pose_resync": 0} if args.no_resync else {}))
    trk = api.Tracker(cfg)
    seq = build_frames(args, dev, rank * T)
    x0 = np.zeros((T, 13)); x0[:, 6:] = seq.pose[0].numpy()
    trk.init(x0)
    pose_np = seq.pose.numpy(); pose_valid_np = seq.pose_valid.numpy().astype(np.uint8)

    def frame_of(step):  # 0, then 1..F cycled
        return 0 if step == 0 else 1 + (step - 1) % F

    def stale(step):  # DatasetImageSegmentationDelayed.cpp:42-63: frame delivered (late) at this step, or None
        idx = step - D
        if idx % D != 0 or (args.single_mask and step > 0):
            return None
        return frame_of(max(idx, 0))

    def do_step(step, host=None):
        f = frame_of(step)
        s = stale(step)
        pose = pose_np[s] if s is not None else None
        pv = pose_valid_np[s] if s is not None else None
        if host is None:
            trk.step(seq.depth[f], seq.flow[f] if step > 0 else None, seq.mask[s] if s is not None else None,
                     pose=pose, pose_valid=pv, device=True)
        else:
            hd, hf, hm = host
            trk.step(hd[f % len(hd)], hf[f % len(hf)] if step > 0 else None, hm[0] if s is not None else None,
                     pose=pose, pose_valid=pv, device=False)

    def barrier():
        torch.cuda.synchronize()
        trk.sync()
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput ------------------------------------------------------
    step = 0
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        do_step(step); step += 1
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
    for _ in range(args.steps):
        do_step(step); step += 1
        host_t.append(time.perf_counter())
        if args.per_step:  # diagnostic: main-stream time stamps per step (velocity chain only)
            e = torch.cuda.Event(enable_timing=True); e.record(ext); step_ev.append((step - 1, e))
    host_issue_ms = (host_t[-1] - host_t[0]) * 1e3 / args.steps
    trk.join()
    ev1.record(ext)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    launches = trk.kernel_launches - l0
    phases, psteps = trk.profile(False)
    if args.per_step and rank == 0:
        prev = ev0
        per = {}
        for st_i, e in step_ev:
            per.setdefault((st_i - D) % D, []).append(prev.elapsed_time(e)); prev = e
        print("per-step ms by (step-D)%D:", {k: round(sum(v) / len(v), 3) for k, v in sorted(per.items())}, file=sys.stderr)
        hper = {}
        for i, (st_i, _) in enumerate(step_ev):
            hper.setdefault((st_i - D) % D, []).append((host_t[i + 1] - host_t[i]) * 1e3)
        print("host issue ms by (step-D)%D:", {k: round(sum(v) / len(v), 3) for k, v in sorted(hper.items())}, file=sys.stderr)
    clocks = sampler.finish(t0, t0 + wall) if sampler else None
    tms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total = float(tms.item())
    ms_per_step = ms_total / args.steps
    value = world * T * args.steps / (ms_total * 1e-3)

    # sanity: the tracker must actually track (guards against a silently skipped pipeline)
    pm, vm = trk.state()
    cnt, _, _ = trk.velocity_info()

    # ---- end to end through the C ABI with host buffers -----------------------------------
    e2e = None
    if not args.no_e2e:
        nh = 2
        hd = [torch.empty((T, H, W), dtype=torch.float32).pin_memory() for _ in range(nh)]
        hf = [torch.empty((T, H, W, 2), dtype=torch.float32).pin_memory() for _ in range(nh)]
        hm = [torch.empty((T, H, W), dtype=torch.uint8).pin_memory()]
        for i in range(nh):
            hd[i].copy_(seq.depth[1 + i]); hf[i].copy_(seq.flow[1 + i])
        hm[0].copy_(seq.mask[1])
        hdn = [x.numpy() for x in hd]; hfn = [x.numpy() for x in hf]; hmn = [x.numpy() for x in hm]
        trk.init(x0)
        st = 0
        do_step(st, (hdn, hfn, hmn)); st += 1
        do_step(st, (hdn, hfn, hmn)); st += 1
        trk.state()
        barrier()
        t0 = time.perf_counter()
        h2d = 0
        for _ in range(args.e2e_steps):
            do_step(st, (hdn, hfn, hmn))
            h2d += T * (H * W * 4 + H * W * 8) + (T * H * W if stale(st) is not None else 0)
            st += 1
            trk.state()  # device->host read of the step's result (pose 13 + velocity 6 doubles per track)
        barrier()
        e2e_s = time.perf_counter() - t0
        ts = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e = {"value": world * T * args.e2e_steps / float(ts.item()), "unit": "tracked frames/s",
               "h2d_bytes_per_step": world * h2d // args.e2e_steps, "d2h_bytes_per_step": world * T * 19 * 8,
               "steps": args.e2e_steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = measured_peak_gbs()
    # Roofline.  The unit of SURVEY 8(d) is a track-frame (depth + flow + mask, each input byte counted once) and it is
    # consumed by the kernel SEQUENCE of a step, so `achieved` = algorithmic bytes of the step / device time of the
    # step.  The dominant kernel is reported beside it with the bytes IT has to touch: the worklist restricts both
    # streaming passes to the non-empty 128-px units of the mask (pass A: mask + depth + flow in, norms out; pass B:
    # depth + flow + norms in), so quoting the whole-frame figure against one pass would exceed the peak.
    dom = max(("flow_pass_a", "flow_pass_b"), key=lambda k: phases[k])
    dom_ms = phases[dom]
    units, _ = trk.worklist()
    unit_bytes = {"flow_pass_a": 128 * (1 + 4 + 8 + 4), "flow_pass_b": 128 * (4 + 8 + 4)}[dom]
    dom_bytes = int(units.astype(np.int64).sum()) * unit_bytes
    dom_gbs = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else None
    step_bytes_gbs = T * BYTES_PER_TRACK_FRAME / (ms_per_step * 1e-3) / 1e9
    achieved = step_bytes_gbs
    dom_traffic = ncu_traffic(dom, T, args.accum != "fp64")
    out = {
        "metric": "tracked frames/sec at 1280x720 (batched tracks)", "value": value, "unit": "tracked frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"auto": "f32 per-pixel terms + f64 reduction/solve (f64 terms for tracks < 32768 px)", "fp64": "f64", "fp32": "f32"}[args.accum],
        "data": "synthetic",
        "config": {"workload": workload_name(args), "tracks_per_gpu": T, "resident_frames": F,
                   "l2_policy": f"inputs larger than L2: {T * BYTES_PER_TRACK_FRAME / 1e9:.2f} GB touched per step, no flush needed",
                   "accumulation": args.accum, "parallelism": f"tracks partitioned over {world} GPU(s), no collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("step", T, args.accum != "fp64"), "peak_kind": peak_kind,
                     "scope": "whole step: algorithmic bytes of T track-frames / device time of the step (all kernels, all streams)",
                     "algorithmic_bytes_per_track_frame": BYTES_PER_TRACK_FRAME,
                     "algorithmic_bytes_per_step": T * BYTES_PER_TRACK_FRAME,
                     "dominant_kernel": {"name": dom, "ms": dom_ms, "touched_units_mean": float(units.mean()),
                                         "algorithmic_bytes": dom_bytes, "achieved": dom_gbs,
                                         "frac": (dom_gbs / peak) if dom_gbs else None, "traffic": dom_traffic,
                                         "note": "duration from CUDA events inside the overlapped step (other streams "
                                                 "share the SMs); profiles/ holds the isolated ncu duration"}},
        "phases_ms_per_step": phases,
        "gpu_launches": int(launches),
        "host_issue_ms_per_step": round(host_issue_ms, 4),
        "clocks": clocks,
        "e2e": e2e,
        "wall_s": wall,
        "sanity": {"mean_valid_pixels": float(cnt.mean()), "mean_abs_w": float(np.abs(vm[:, 3:]).mean()),
                   "finite": bool(np.isfinite(pm).all() and np.isfinite(vm).all())},
    }
    if not args.no_cpu and world == 1:  # reported at N = 1 only (rank 0)
        try:
            out["cpu_baseline"] = cpu_baseline(args, seq, "port")
        except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
            out["cpu_baseline"] = {"error": repr(e)}
    if world == 1 and not args.no_sweep and not args.single_mask and T == 256:
        # Same workload at twice the batch: the per-step latencies that do not scale with the batch (pose re-sync replay,
        # new-mask scatter chain, launch tails) amortise - reported beside the headline, never instead of it.
        try:
            del trk, seq
            torch.cuda.empty_cache()
            out["batch_sweep"] = [resident_throughput(args, api, dev, local, 512, 6, 60, 12, peak)]
        except Exception as e:
            out["batch_sweep"] = {"error": repr(e)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def resident_throughput(args, api, dev, local, T, F, steps, warmup, peak):
    """Device-resident throughput of the default workload at another batch size (single GPU)."""
    import numpy as np
    import torch
    from roft_b200.synthetic import make_sequence
    D = args.delay
    cfg = api.default_config(n_tracks=T, subsampling_radius=args.stride, segm_delay=D, pose_delay=D, device=local,
                             accum_fp64={"fp32": 0, "fp64": 1, "auto": 2}[args.accum])
    trk = api.Tracker(cfg)
    seq = make_sequence(T, F + 1, W, H, device=dev, target_coverage=args.coverage, first_track_id=0, track_chunk=8)
    torch.cuda.synchronize()
    x0 = np.zeros((T, 13)); x0[:, 6:] = seq.pose[0].numpy()
    trk.init(x0)
    pose_np = seq.pose.numpy(); pv_np = seq.pose_valid.numpy().astype(np.uint8)

    def do_step(step):
        f = 0 if step == 0 else 1 + (step - 1) % F
        idx = step - D
        s = None if idx % D != 0 else (0 if idx <= 0 else 1 + (idx - 1) % F)
        trk.step(seq.depth[f], seq.flow[f] if step > 0 else None, seq.mask[s] if s is not None else None,
                 pose=pose_np[s] if s is not None else None, pose_valid=pv_np[s] if s is not None else None, device=True)

    step = 0
    for _ in range(warmup):
        do_step(step); step += 1
    torch.cuda.synchronize(); trk.sync()
    ext = torch.cuda.ExternalStream(trk.stream, device=dev)
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    for _ in range(steps):
        do_step(step); step += 1
    trk.join()
    ev1.record(ext)
    torch.cuda.synchronize(); trk.sync()
    ms = ev0.elapsed_time(ev1) / steps
    gbs = T * BYTES_PER_TRACK_FRAME / (ms * 1e-3) / 1e9
    pm, vm = trk.state()
    return {"tracks_per_gpu": T, "resident_frames": F, "steps": steps, "warmup": warmup, "value": T / (ms * 1e-3),
            "unit": "tracked frames/s", "ms_per_step": ms, "roofline_achieved": gbs, "roofline_frac": gbs / peak,
            "finite": bool(np.isfinite(pm).all() and np.isfinite(vm).all())}


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
    seq = make_sequence(n_tracks, min(args.frames, 6) + 1, W, H, device="cpu", target_coverage=args.coverage, track_chunk=4)
    t0 = time.perf_counter()
    cb = cpu_baseline(args, seq, "port", threads=ncpu, seconds=max(10.0, args.cpu_seconds))
    out = {
        "impl": "reference", "metric": "tracked frames/sec at 1280x720 (batched tracks)", "value": cb["value"],
        "unit": "tracked frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * n_tracks / cb["value"] if cb["value"] else None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "note": "CPU restatement of the reference (the reference itself needs Eigen/OpenCV/bfl, absent here); "
                   "one thread per track over all host cores"},
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
