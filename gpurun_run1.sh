mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -x 2>&1 | tail -5
for r in 1 0; do
export ROFTB_RING=$r
timeout 300 python bench.py --no-cpu --no-e2e --steps 48 --warmup 12 > gpurun_out/b.log 2>&1
tail -1 gpurun_out/b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ring $r', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['sanity'])"
done
