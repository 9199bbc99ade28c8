timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
for extra in "" "--fp32-accum"; do
timeout 800 python bench.py --no-cpu --no-e2e $extra --steps 12 --warmup 4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})"
done
ncu --set full --clock-control none --import-source on -k regex:"k_flow_pass|k_warp_scatter" -s 14 -c 6 -o gpurun_out/prof_r1c python bench.py --no-cpu --no-e2e --tracks 256 --frames 7 --steps 3 --warmup 2 > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log | cut -c1-100
