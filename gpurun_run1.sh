mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --no-e2e --steps 48 --warmup 12 > gpurun_out/b.log 2>&1
tail -1 gpurun_out/b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['roofline'], d['clocks'])"
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_sel" -s 30 -c 6 --csv --log-file gpurun_out/passes.csv python bench.py --no-cpu --no-e2e --steps 12 --warmup 12 > gpurun_out/ncu_k.log 2>&1
python - <<'PY'
import csv,re
rows=[r for r in csv.reader(l for l in open('gpurun_out/passes.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    m=re.search(r'(k_\w+)(<[^>]*>)?',r[ki]); d.setdefault((int(r[ii]),m.group(0)[:30]),{})[r[mi].split('.')[0][-18:]]=r[vi]
for k,v in sorted(d.items()): print(k, v)
PY
