#!/bin/bash
# circular / GlowBase tests, launch list of the bench steps, train-step A/B against the round-2a library
mkdir -p gpurun_out
T=r03a
( time python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > gpurun_out/${T}_pytest.log 2>&1
tail -9 gpurun_out/${T}_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_rqs|diag_gauss|sum_stage|fill_kernel" -c 200 --csv --log-file gpurun_out/${T}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs > gpurun_out/${T}_ncu_launch.log 2>&1
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-eager --no-extra-configs"
for v in new r2a new r2a; do
  if [ $v = r2a ]; then export NFB200_LIB=$PWD/tools/ab/libnfb200_r2a.so; else unset NFB200_LIB; fi
  $B > gpurun_out/${T}_train_$v.json 2>> gpurun_out/${T}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_train_$v.json').read().strip().splitlines()[-1])
print('$v', 'fwd ms/step', round(d['ms_per_step'],3), 'train_step ms', round(d['train_step']['ms_per_step'],2))
PY
done
unset NFB200_LIB
python tools/train_step_probe.py 2>&1 | grep step
tail -3 gpurun_out/${T}.err
