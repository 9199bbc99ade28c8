python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r1_2gpu.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_2gpu.json')); print(round(d['value']), d['n_gpus'], round(d['ms_per_step'],3), d['e2e'], d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --cpu-seconds 5 2>&1 | tail -1 | cut -c1-200
