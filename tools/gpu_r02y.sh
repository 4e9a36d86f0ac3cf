#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02y}
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs"
run() { name=$1; shift; env "$@" $B > gpurun_out/${T}_$name.json 2>> gpurun_out/${T}.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'loss', d['config']['loss'])
except Exception as e: print('$name', 'FAILED', e)
PY
}
run A_f10 X=1
run B_f8 NFB_FPC=8
run C_f6 NFB_FPC=6
run D_f10 X=2
run E_f8 NFB_FPC=8
NFB_FPC=8 python -m pytest tests -m gpu -q -x -k "bench_config or trained or log_prob_and_kld or full_batch" 2>&1 | tail -3
NFB_FPC=8 python tools/gpu_debug.py prof 65536 > gpurun_out/${T}_prof_f8.log 2>&1
grep -A1 "abs  :" gpurun_out/${T}_prof_f8.log | cut -c1-800
tail -3 gpurun_out/${T}.err
