#!/bin/bash
mkdir -p gpurun_out
T=r03f
( python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs"
run() { name=$1; shift; env "$@" $B > gpurun_out/${T}_$name.json 2>> gpurun_out/${T}.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'loss', d['config']['loss'])
except Exception as e: print('$name', 'FAILED', e)
PY
}
run A_tickets X=1
run B_static NFB_STATIC_UNITS=1
run C_prev NFB200_LIB=$PWD/tools/ab/libnfb200_prev.so
run D_tickets X=2
run E_coupled_tickets NFB_BENCH_KIND=coupled
run F_coupled_static NFB_BENCH_KIND=coupled NFB_STATIC_UNITS=1
python tools/train_step_probe.py 2>&1 | grep -E "repack|^step 2"
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/${T}_racecheck.log 2>&1; tail -3 gpurun_out/${T}_racecheck.log
tail -3 gpurun_out/${T}.err
