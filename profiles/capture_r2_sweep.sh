#!/bin/bash
# DRAM traffic of the velocity kernel at the other mask coverages (bench.py workloads c5 / full / ref): one --set full launch each
set -u
mkdir -p gpurun_out
export ROFTB_PARTS=1
for wl in c5 full ref; do
  timeout 400 ncu --set full --clock-control none -k regex:"k_velocity_track" -s 14 -c 1 -o /tmp/prof_r02_$wl -f \
      python bench.py --workload $wl --no-cpu --no-e2e --no-extras --no-parity --steps 4 --warmup 12 > gpurun_out/r02_ncu_$wl.log 2>&1
  ncu -i /tmp/prof_r02_$wl.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_velocity_$wl.csv 2>/dev/null
done
ls -la gpurun_out/r02_ncu_full_velocity_*.csv
