#!/bin/bash
mkdir -p gpurun_out
T=r03e
( python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
python tools/train_step_probe.py 2>&1 | grep -E "repack|^step 2|adam step 2"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-eager --no-extra-configs > gpurun_out/${T}_bench.json 2> gpurun_out/${T}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r03e_bench.json').read().strip().splitlines()[-1])
print('fwd ms/step', round(d['ms_per_step'],3), 'train_step ms', round(d['train_step']['ms_per_step'],2))
PY
