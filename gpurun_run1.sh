mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --no-e2e --per-step --steps 48 --warmup 12 > gpurun_out/b.log 2>&1
grep "per-step" gpurun_out/b.log
tail -1 gpurun_out/b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['sanity'])"
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_flow_pass_ring" -s 40 -c 4 --csv --log-file gpurun_out/passes.csv python bench.py --no-cpu --no-e2e --steps 12 --warmup 12 > gpurun_out/ncu_k.log 2>&1
grep -o '"k_flow_pass_ring[^"]*\|gpu__time[^,]*,"[a-z]*","[0-9,]*\|smsp__inst[^,]*,"[a-z]*","[0-9,]*' gpurun_out/passes.csv | paste - - - | sed 's/unnamed>:://' | cut -c1-160
