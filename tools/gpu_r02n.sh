#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02n_bench_2gpu.json 2> gpurun_out/r02n_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02n_ref_2gpu.json 2> gpurun_out/r02n_ref_2gpu.err
tail -3 gpurun_out/r02n_bench_2gpu.err; python -c "
import json
d=json.loads(open('gpurun_out/r02n_bench_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['train_step'])"; tail -c 300 gpurun_out/r02n_ref_2gpu.json
