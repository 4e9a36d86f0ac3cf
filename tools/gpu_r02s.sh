#!/bin/bash
# round 2b v3 kernel: 3 fixed slots, density path without the x tile in shared memory, LU fold
mkdir -p gpurun_out
T=${1:-r02s}
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs"
run() { name=$1; shift; env "$@" $B > gpurun_out/${T}_$name.json 2>> gpurun_out/${T}.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'loss', d['config']['loss'])
except Exception as e: print('$name', 'FAILED', e)
PY
}
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/${T}_pytest.log 2>&1
tail -8 gpurun_out/${T}_pytest.log
run A_default NFB_DEBUG_PACK=1
grep "nfb pack" gpurun_out/${T}.err | sort | uniq -c | head -5
run B_nofold NFB_NO_FOLD=1
NFB_BENCH_KIND=coupled run C_coupled NFB_BENCH_KIND=coupled
python tools/gpu_debug.py prof 65536 > gpurun_out/${T}_prof.log 2>&1
head -c 1800 gpurun_out/${T}_prof.log
grep -v "nfb pack" gpurun_out/${T}.err | tail -5
