mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_flow_pass|k_warp_|k_sel_|k_tile_|k_ukf|k_vel_|k_mask_" -s 170 -c 178 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --no-cpu --no-e2e --no-sweep --steps 12 --warmup 12 > gpurun_out/ncu_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_flow_pass_ring|k_sel_|k_warp_scatter|k_ukf_batch|k_tile_count" -s 98 -c 16 -o gpurun_out/prof_r1_final -f python bench.py --no-cpu --no-e2e --no-sweep --steps 12 --warmup 12 > gpurun_out/ncu_final2.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; tail -c 400 gpurun_out/bench_r1_ref.json
ls -la gpurun_out | tail -4
