#!/bin/bash
# Commands behind the files in profiles/ (run on a B200 box, e.g. `gpurun --timeout 2400 -- 'bash profiles/capture.sh'`).
# Numbers printed by a run under ncu are never bench values; only shares and per-kernel counters are used.
set -u
mkdir -p gpurun_out
K='k_flow_pass|k_warp_|k_sel_|k_tile_|k_ukf|k_vel_|k_mask_'
# launch list (per-launch durations of the repo's own kernels, cold cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -s 170 -c 178 --csv \
    --log-file gpurun_out/launches_r1_final.csv python bench.py --no-cpu --no-e2e --no-sweep --steps 12 --warmup 12 > gpurun_out/ncu_final.log 2>&1
# full capture of the top kernels (source page included)
ncu --set full --clock-control none --import-source on -k regex:"k_flow_pass_ring|k_sel_|k_warp_scatter|k_ukf_batch|k_tile_count" \
    -s 98 -c 16 -o gpurun_out/prof_r1_final -f python bench.py --no-cpu --no-e2e --no-sweep --steps 12 --warmup 12 > gpurun_out/ncu_final2.log 2>&1
# read back here with:  ncu -i gpurun_out/prof_r1_final.ncu-rep --page raw --csv   /   --page source --csv --print-source sass
