#!/bin/bash
# cluster-size / register-cap sweep of the velocity kernel (device-resident bench, no CPU baseline)
for o in ${REGS:-96 128}; do
for c in ${CLUSTERS:-4 8}; do
  echo "== ROFTB_CLUSTER=$c ROFTB_VT_REGS=$o"
  ROFTB_PHASE_DEBUG=1 ROFTB_CLUSTER=$c ROFTB_VT_REGS=$o timeout 300 python bench.py --steps ${STEPS:-24} --warmup 12 --no-cpu --no-sweep --no-e2e $BENCH_ARGS 2> gpurun_out/sweep_c${c}_o$o.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],3),'phases',{k:round(v,3) for k,v in d['phases_ms_per_step'].items()},'valid',d['sanity'])"
  grep roftb gpurun_out/sweep_c${c}_o$o.err
done
done
