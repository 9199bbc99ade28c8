mkdir -p gpurun_out
for cfg in "10 9" "10 13" "10 18" "10 27" "12 18"; do
set -- $cfg
export ROFTB_BPT=$1 ROFTB_SB=$2
timeout 800 python bench.py --no-cpu --no-e2e --steps 48 --warmup 12 > gpurun_out/b.log 2>&1
tail -1 gpurun_out/b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})"
done
