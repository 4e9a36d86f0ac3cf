#!/bin/bash
mkdir -p gpurun_out
T=r03l
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs"
run() { name=$1; shift; env "$@" $B > gpurun_out/${T}_$name.json 2>> gpurun_out/${T}.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'loss', d['config']['loss'])
except Exception as e: print('$name', 'FAILED', e)
PY
}
run A_ar_new X=1
run B_ar_prev NFB200_LIB=$PWD/tools/ab/libnfb200_prev.so
run C_coupled_new NFB_BENCH_KIND=coupled
run D_coupled_prev NFB_BENCH_KIND=coupled NFB200_LIB=$PWD/tools/ab/libnfb200_prev.so
run E_coupled_new NFB_BENCH_KIND=coupled
( python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) > gpurun_out/${T}_pytest.log 2>&1
tail -2 gpurun_out/${T}_pytest.log
