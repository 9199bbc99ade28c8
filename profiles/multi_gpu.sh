#!/bin/bash
# torchrun launch of bench.py on N GPUs of one box, as the driver does: `gpurun --gpus N -- "bash profiles/multi_gpu.sh N"`
mkdir -p gpurun_out
N=$1
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-60} --warmup 12 --no-cpu $BENCH_ARGS > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
echo "rc=$?"
wc -c gpurun_out/bench_r2_${N}gpu.json gpurun_out/bench_r2_${N}gpu.err
tail -5 gpurun_out/bench_r2_${N}gpu.err | cut -c1-400
tail -c 1200 gpurun_out/bench_r2_${N}gpu.json
