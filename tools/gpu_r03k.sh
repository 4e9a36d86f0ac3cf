#!/bin/bash
mkdir -p gpurun_out
T=r03k
( python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs"
run() { name=$1; shift; env "$@" $B > gpurun_out/${T}_$name.json 2>> gpurun_out/${T}.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'e2e loss', d['e2e']['loss'])
except Exception as e: print('$name', 'FAILED', e)
PY
}
run A_new X=1
run B_nowave NFB_NO_WAVE_ORDER=1
run C_new X=2
run D_nowave NFB_NO_WAVE_ORDER=1
tail -3 gpurun_out/${T}.err
