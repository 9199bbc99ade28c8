mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ukf_batch|k_flow_pass" -s 60 -c 10 -o gpurun_out/prof_r1j -f python bench.py --no-cpu --no-e2e --steps 12 --warmup 12 > gpurun_out/ncu_j.log 2>&1
tail -3 gpurun_out/ncu_j.log
