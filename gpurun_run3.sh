mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_flow_pass" -s 60 -c 12 --csv --log-file gpurun_out/passes.csv python bench.py --no-cpu --no-e2e --steps 12 --warmup 12 > gpurun_out/ncu_k.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/passes.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][r[ki].find('k_flow_pass'):][:32]),{})[r[mi]]=r[vi]
for k,v in d.items(): print(k, v)
PY
