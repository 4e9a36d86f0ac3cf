#!/bin/bash
# round 2b: early chunk epilogue x merged records A/B (production build without profiling stamps)
mkdir -p gpurun_out
T=${1:-r02v}
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs"
run() { name=$1; shift; env "$@" $B > gpurun_out/${T}_$name.json 2>> gpurun_out/${T}.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'loss', d['config']['loss'])
except Exception as e: print('$name', 'FAILED', e)
PY
}
run A_default X=1
run B_noearly NFB_NO_EARLY_EPI=1
run C_nomerge NFB_NO_MERGE=1
run D_neither NFB_NO_EARLY_EPI=1 NFB_NO_MERGE=1
run E_coupled_default NFB_BENCH_KIND=coupled
run F_coupled_neither NFB_BENCH_KIND=coupled NFB_NO_EARLY_EPI=1 NFB_NO_MERGE=1
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
python tools/gpu_debug.py prof 65536 > gpurun_out/${T}_prof.log 2>&1
grep -A1 "abs  :" gpurun_out/${T}_prof.log | cut -c1-700
tail -3 gpurun_out/${T}.err
