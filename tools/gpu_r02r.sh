#!/bin/bash
# round 2b experiment matrix: record format x MMA order x LU fold x tile prefetch
mkdir -p gpurun_out
T=r02r
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs"
run() { name=$1; shift; env "$@" $B > gpurun_out/${T}_$name.json 2>> gpurun_out/${T}.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'loss', d['config']['loss'])
except Exception as e: print('$name', 'FAILED', e)
PY
}
run A_default NFB_X=1
run B_order1 NFB_MMA_ORDER=1
run C_order2 NFB_MMA_ORDER=2
run D_plain NFB_PLAIN_RECORDS=1
run E_plain_nofold NFB_PLAIN_RECORDS=1 NFB_NO_FOLD=1
run F_plain_nofold_nopre NFB_PLAIN_RECORDS=1 NFB_NO_FOLD=1 NFB_NO_PREFETCH=1
run G_mixed_nofold NFB_NO_FOLD=1
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/${T}_pytest.log 2>&1
tail -6 gpurun_out/${T}_pytest.log
( NFB_PLAIN_RECORDS=1 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/${T}_pytest_plain.log 2>&1
tail -4 gpurun_out/${T}_pytest_plain.log
python tools/gpu_debug.py prof 65536 > gpurun_out/${T}_prof_default.log 2>&1
NFB_PLAIN_RECORDS=1 python tools/gpu_debug.py prof 65536 > gpurun_out/${T}_prof_plain.log 2>&1
tail -5 gpurun_out/${T}.err
