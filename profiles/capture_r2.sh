#!/bin/bash
# Round-2 commands behind the files in profiles/ (run on a B200 box: `gpurun --timeout 1500 -- 'bash profiles/capture_r2.sh'`).
# Numbers printed by a run under ncu are never bench values; only shares and per-kernel counters are used.
set -u
mkdir -p gpurun_out
# (ROFTB_PARTS=1: one launch per step over all 256 tracks; capture_r2_parts.sh holds the launch list of the default, two part contexts)
export ROFTB_PARTS=1
B="python bench.py --no-cpu --no-e2e --no-sweep --no-parity --steps 12 --warmup 12"
# launch list: per-launch durations of the repo's own kernels (cold cache, serialised), two mask periods
if [ "${1:-all}" != "full" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -s 100 -c 160 --csv \
    --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_list.log 2>&1
fi
# full capture (source page included): two launches of the velocity kernel; the new-mask scatter and the pose UKF of an event step
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_velocity_track" -s 14 -c 2 \
    -o /tmp/prof_r02_velocity -f $B > gpurun_out/r02_ncu_full_velocity.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_warp_scatter|k_ukf_batch" -s 24 -c 16 \
    -o /tmp/prof_r02_event -f $B > gpurun_out/r02_ncu_full_event.log 2>&1
for n in velocity event; do
  ncu -i /tmp/prof_r02_$n.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_$n.csv 2>/dev/null
done
ncu -i /tmp/prof_r02_velocity.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_ncu_source_velocity.csv 2>/dev/null
ls -la /tmp/prof_r02_*.ncu-rep
# keep the velocity report itself if it fits the 64 MiB limit of gpurun_out/
sz=$(stat -c %s /tmp/prof_r02_velocity.ncu-rep); [ "$sz" -lt 30000000 ] && cp /tmp/prof_r02_velocity.ncu-rep gpurun_out/
