#!/bin/bash
mkdir -p gpurun_out
T=r03m
( python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > gpurun_out/${T}_pytest.log 2>&1
tail -2 gpurun_out/${T}_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_rqs|diag_gauss|sum_stage|fill_kernel" -c 200 --csv --log-file gpurun_out/${T}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs > gpurun_out/${T}_ncu_launch.log 2>&1
grep -c fused_rqs gpurun_out/${T}_launches_bench_steps2.csv; grep -c nan gpurun_out/${T}_launches_bench_steps2.csv
CUDA_LAUNCH_BLOCKING=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs 2>&1 | tail -1 | cut -c1-200
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
