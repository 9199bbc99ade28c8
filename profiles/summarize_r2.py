#!/usr/bin/env python3
"""Summaries of the round-2 ncu captures (profiles/capture_r2.sh) -> profiles/r02_*.txt.

  python profiles/summarize_r2.py gpurun_out

* r02_launches_256tracks.txt     per-kernel launch durations (cold cache, serialised) and their share of a mask period
* r02_ncu_full_summary.txt       headline counters of the velocity kernel, the new-mask scatter and the pose UKF
* r02_ncu_source_summary.txt     the velocity kernel's SASS split at its cluster barriers: instructions and stall samples per phase
"""
import collections
import csv
import os
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
out = os.path.dirname(os.path.abspath(__file__))


def short(name):
    return name.split("(")[0].replace("void ", "").replace("roftb::<unnamed>::", "").replace("unnamed>::", "")


def launches(name="r02_launches", parts=1, out_name="r02_launches_256tracks.txt", note=""):
    rows = list(csv.reader(open(os.path.join(src, name + ".csv" if os.path.exists(os.path.join(src, name + ".csv")) else name + "_256tracks.csv"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    seq = []
    for r in rows[hi + 1:]:
        if len(r) > vi and r[0].isdigit():
            v = float(r[vi].replace(",", ""))
            v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
            seq.append((short(r[ki]), v))
    # one mask period = from one k_tile_count (a delivery) to the next
    starts = [i for i, (n, _) in enumerate(seq) if n == "k_tile_count"]
    period = seq[starts[0]:starts[parts]] if len(starts) > parts else seq  # (every part context delivers in the same step)
    d = collections.OrderedDict()
    for n, v in period:
        d.setdefault(n, []).append(v)
    tot = sum(v for _, v in period)
    n_steps = sum(1 for n, _ in period if n.startswith("k_velocity_track")) // parts
    with open(os.path.join(out, out_name), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none (profiles/capture_r2.sh), bench.py default workload, 256 tracks" + note + "\n")
        f.write(f"one mask period = {n_steps} steps, {len(period)} launches, {tot:.1f} us serialised and cold = {tot / n_steps:.1f} us per step\n\n")
        f.write(f"{'kernel':36s} {'launches':>8s} {'mean us':>10s} {'min':>9s} {'max':>9s} {'share':>7s}\n")
        for n, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{n:36s} {len(v):8d} {sum(v) / len(v):10.1f} {min(v):9.1f} {max(v):9.1f} {100 * sum(v) / tot:6.1f}%\n")
        f.write("\nlaunch sequence of the period (us):\n")
        for n, v in period:
            f.write(f"  {n:36s} {v:9.1f}\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem"]


def full():
    with open(os.path.join(out, "r02_ncu_full_summary.txt"), "w") as f:
        f.write("ncu --set full --clock-control none (profiles/capture_r2.sh); per launch, cold cache, kernel alone on the GPU\n")
        for fn in ("r02_ncu_full_velocity.csv", "r02_ncu_full_event.csv"):
            rows = list(csv.reader(open(os.path.join(src, fn))))
            h, units = rows[0], rows[1]
            stall = [i for i, c in enumerate(h) if c.startswith("smsp__average_warps_issue_stalled") and c.endswith("per_issue_active.ratio")]
            for r in rows[2:]:
                dur = float(r[h.index("gpu__time_duration.sum")])
                if dur < 0.1 or (units[h.index("gpu__time_duration.sum")] == "us" and dur < 100):
                    continue  # the idle launches of steps without a delivery
                f.write(f"\n{short(r[h.index('Kernel Name')])}\n")
                for w in WANT:
                    if w in h:
                        f.write(f"  {w:64s} {r[h.index(w)]:>16s} {units[h.index(w)]}\n")
                top = sorted(((float(r[i] or 0), h[i]) for i in stall), reverse=True)[:5]
                f.write("  top stalls (warps per issue-active cycle): " +
                        ", ".join(f"{c.split('issue_stalled_')[1].split('_per_')[0]} {v:.2f}" for v, c in top) + "\n")


def source():
    rd = csv.reader(open(os.path.join(src, "r02_ncu_source_velocity.csv")))
    hdr, rows = None, []
    for r in rd:
        if r and r[0] == "Address":
            hdr = r
        elif hdr and r and r[0].startswith("0x"):
            rows.append(r)
    ia, ii, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stalls = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    # the report holds the launches back to back: keep the first (addresses restart)
    first = rows[0][0]
    n1 = next((k for k in range(1, len(rows)) if rows[k][0] == first), len(rows))
    rows = rows[:n1]
    names = ["prologue (flags -> worklist, lazy clear, slot)", "pass A (ring, gates, innovations, fused scatter, records)",
             "pairing + level-0 histogram", "level 1", "level 2 + statistics (+ finish)", "pass B (40 sums)", "partials / release"]
    seg, cur = [], []
    for r in rows:
        cur.append(r)
        if "UCGABAR_WAIT" in r[ia]:
            seg.append(cur)
            cur = []
    seg.append(cur)
    tot = sum(int(r[ii] or 0) for r in rows)
    tots = sum(int(r[isamp] or 0) for r in rows)
    with open(os.path.join(out, "r02_ncu_source_summary.txt"), "w") as f:
        f.write("k_velocity_track<1,128>: SASS of one launch (256 tracks) split at the cluster barriers (UCGABAR_WAIT)\n")
        f.write(f"{len(rows)} SASS instructions, {tot / 1e6:.1f} M warp-instructions executed, {tots} stall samples\n\n")
        f.write(f"{'segment':62s} {'SASS':>6s} {'M instr':>9s} {'share':>7s} {'samples':>8s}  top stall reasons\n")
        for k, s in enumerate(seg):
            n = sum(int(r[ii] or 0) for r in s)
            sm = sum(int(r[isamp] or 0) for r in s)
            st = sorted(((sum(int(r[i] or 0) for r in s), hdr[i][6:]) for i in stalls), reverse=True)[:4]
            nm = names[k] if k < len(names) else f"segment {k}"
            f.write(f"{nm:62s} {len(s):6d} {n / 1e6:9.1f} {100 * n / max(tot, 1):6.1f}% {100 * sm / max(tots, 1):7.1f}%  " +
                    ", ".join(f"{c} {100 * v / max(sm, 1):.0f}%" for v, c in st) + "\n")
        mn = collections.Counter()
        for r in rows:
            op = r[ia].split()[0] if not r[ia].strip().startswith("@") else r[ia].split()[1]
            mn[op.split(".")[0]] += int(r[ii] or 0)
        f.write("\nmost executed opcodes (M warp-instructions): " + ", ".join(f"{o} {v / 1e6:.1f}" for o, v in mn.most_common(14)) + "\n")
        keys = ("UBLKCP", "LDGSTS", "SYNCS", "UCGABAR", "FFMA2", "REDS", "ATOMS", "LDG", "STG")
        f.write("static counts: " + ", ".join(f"{k} {sum(1 for r in rows if k in r[ia])}" for k in keys) + "\n")


if __name__ == "__main__":
    if os.path.exists(os.path.join(src, "r02_launches.csv")) or os.path.exists(os.path.join(src, "r02_launches_256tracks.csv")):
        launches()
        full()
        source()
    if os.path.exists(os.path.join(src, "r02_launches_parts2.csv")) or os.path.exists(os.path.join(src, "r02_launches_parts2_256tracks.csv")):
        launches("r02_launches_parts2", 2, "r02_launches_parts2_256tracks.txt",
                 "; default build: two pipelined part contexts of 128 tracks, every kernel launched once per part and step")
