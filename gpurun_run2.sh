mkdir -p gpurun_out
for pa in 0 1 2 0 2; do
export ROFTB_PREP_AFTER=$pa
timeout 300 python bench.py --no-cpu --no-e2e --per-step --steps 96 --warmup 12 > gpurun_out/b.log 2>&1
tail -1 gpurun_out/b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('prep_after $pa', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})"
done
