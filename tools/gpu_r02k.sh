#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -x 2>&1 | tail -30 > gpurun_out/r02k_pytest.log
python tools/train_step_probe.py > gpurun_out/r02k_train_probe.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_rqs|diag_gauss|sum_stage|fill_kernel" -c 400 --csv --log-file gpurun_out/r02k_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs > gpurun_out/r02k_ncu_launch.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-eager > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
tail -12 gpurun_out/r02k_pytest.log; grep step gpurun_out/r02k_train_probe.log; python -c "
import json
d=json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['train_step'])"
