mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_sel|k_flow_pass_ring" -s 120 -c 24 --csv --log-file gpurun_out/passes.csv python bench.py --no-cpu --no-e2e --steps 12 --warmup 12 > gpurun_out/ncu_k.log 2>&1
python - <<'PY'
import csv,re
rows=[r for r in csv.reader(l for l in open('gpurun_out/passes.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    m=re.search(r'(k_\w+)(<[^>]*>)?',r[ki]); d.setdefault((int(r[ii]),m.group(0)[:30]),{})[r[mi].split('.')[0][-18:]]=r[vi]
for k,v in sorted(d.items()): print(k, v)
PY
