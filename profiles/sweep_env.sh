#!/bin/bash
# A/B runs of the device-resident bench under different ROFTB_* tuning variables: each argument is one "VAR=val VAR=val" set
for cfg in "$@"; do
  echo "== $cfg"
  env ROFTB_PHASE_DEBUG=1 $cfg timeout 300 python bench.py --steps ${STEPS:-36} --warmup 12 --no-cpu --no-sweep --no-e2e --per-step $BENCH_ARGS 2> gpurun_out/sweep_env.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],3),'phases',{k:round(v,3) for k,v in d['phases_ms_per_step'].items()})"
  grep -E "roftb|per-step" gpurun_out/sweep_env.err
done
