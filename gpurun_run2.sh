mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --no-cpu --no-e2e --no-sweep --per-step --steps 96 --warmup 12 > gpurun_out/b.log 2>&1
tail -1 gpurun_out/b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['sanity'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_sel" -s 30 -c 6 --csv --log-file gpurun_out/passes.csv python bench.py --no-cpu --no-e2e --no-sweep --steps 12 --warmup 12 > gpurun_out/ncu_k.log 2>&1
grep -o 'k_sel_[a-z0-9_]*\|gpu__time[^,]*,"[a-z]*","[0-9,]*\|smsp__inst[^,]*,"[a-z]*","[0-9,]*' gpurun_out/passes.csv | paste - - - - | cut -c1-150
