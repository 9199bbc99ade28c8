mkdir -p gpurun_out
for extra in "" "--no-resync"; do
timeout 800 python bench.py --no-cpu --no-e2e --per-step $extra --steps 48 --warmup 12 > gpurun_out/b.log 2>&1
tail -5 gpurun_out/b.log | cut -c1-300
grep -o '"host_issue_ms_per_step": [0-9.]*' gpurun_out/b.log
done
