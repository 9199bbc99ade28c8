mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/bench_r1_own.json 2> gpurun_out/bench_r1_own.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_own.json').read().strip().split('\n')[-1])
print(round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks'], d['e2e'], d.get('batch_sweep'), d['cpu_baseline']['value'], d['gpu_launches'])
print(d['roofline']['dominant_kernel'])
PY
tail -3 gpurun_out/bench_r1_own.err
