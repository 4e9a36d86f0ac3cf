#!/bin/bash
# round 2b kernel: mixed records + byte ring + tail prefetch + epilogue ld pipelining
mkdir -p gpurun_out
T=${1:-r02q}
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/${T}_pytest.log 2>&1
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step"
$B > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python tools/gpu_debug.py prof 65536 > gpurun_out/${T}_prof.log 2>&1
tail -12 gpurun_out/${T}_pytest.log
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))
for e in d.get('extra', []): print(e.get('config','')[:40], e.get('ms', e.get('ms_exact_2x2')))
PY
tail -3 gpurun_out/${T}_bench.err
head -c 3000 gpurun_out/${T}_prof.log
