#!/bin/bash
# A/B on one box: current build vs the previous commit's library (tools/ab/libnfb200_prev.so)
mkdir -p gpurun_out
T=${1:-r02w}
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs"
run() { name=$1; shift; env "$@" $B > gpurun_out/${T}_$name.json 2>> gpurun_out/${T}.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'loss', d['config']['loss'])
except Exception as e: print('$name', 'FAILED', e)
PY
}
run A_new X=1
run B_prev NFB200_LIB=$PWD/tools/ab/libnfb200_prev.so
run C_new_coupled NFB_BENCH_KIND=coupled
run D_prev_coupled NFB_BENCH_KIND=coupled NFB200_LIB=$PWD/tools/ab/libnfb200_prev.so
run E_new X=2
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/${T}_pytest.log 2>&1
head -4 gpurun_out/${T}_pytest.log | tail -2
python tools/gpu_debug.py prof 65536 > gpurun_out/${T}_prof.log 2>&1
grep -A1 "abs  :" gpurun_out/${T}_prof.log | cut -c1-700
tail -3 gpurun_out/${T}.err
