#!/bin/bash
# Launch list of the DEFAULT bench command (two pipelined part contexts per GPU): every kernel appears once per part and step
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -s 200 -c 320 --csv \
    --log-file gpurun_out/r02_launches_parts2.csv python bench.py --no-cpu --no-e2e --no-sweep --no-parity --steps 12 --warmup 12 \
    > gpurun_out/r02_ncu_list_parts2.log 2>&1
ls -la gpurun_out/r02_launches_parts2.csv
