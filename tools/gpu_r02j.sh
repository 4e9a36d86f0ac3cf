#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step"
$B > gpurun_out/r02j_single_all.json 2> gpurun_out/r02j.err
NFB_POLL_LANE0=1 $B > gpurun_out/r02j_single_lane0_nosleep.json 2>> gpurun_out/r02j.err
for f in single_all single_lane0_nosleep; do python -c "
import json
d=json.loads(open('gpurun_out/r02j_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['e2e']['ms_per_step'],3))"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02j_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step > gpurun_out/r02j_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_rqs -s 4 -c 1 -o gpurun_out/prof_r02 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step > gpurun_out/r02j_ncu_full.log 2>&1
NFB_PROBE_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02j_train_launches.csv python tools/train_step_probe.py > gpurun_out/r02j_ncu_train.log 2>&1
ls -la gpurun_out/prof_r02.ncu-rep
