mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_flow_pass_ring" -s 40 -c 4 -o gpurun_out/prof_r1k -f python bench.py --no-cpu --no-e2e --steps 12 --warmup 12 > gpurun_out/ncu_k.log 2>&1
tail -2 gpurun_out/ncu_k.log | cut -c1-200
