timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -4
for extra in "" "--accum fp64" "--stride 35"; do
timeout 800 python bench.py --no-cpu --no-e2e $extra --steps 24 --warmup 6 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})"
done
