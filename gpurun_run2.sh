mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --no-cpu --no-e2e --no-sweep --per-step --steps 96 --warmup 12 > gpurun_out/b.log 2>&1
grep "per-step" gpurun_out/b.log
tail -1 gpurun_out/b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})"
done
