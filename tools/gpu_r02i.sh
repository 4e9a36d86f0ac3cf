#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step"
$B > gpurun_out/r02i_pair_lane0.json 2> gpurun_out/r02i.err
NFB_POLL_ALL=1 $B > gpurun_out/r02i_pair_all.json 2>> gpurun_out/r02i.err
NFB_NO_PAIR=1 $B > gpurun_out/r02i_nopair_lane0.json 2>> gpurun_out/r02i.err
NFB_NO_PAIR=1 NFB_POLL_ALL=1 $B > gpurun_out/r02i_nopair_all.json 2>> gpurun_out/r02i.err
for f in pair_lane0 pair_all nopair_lane0 nopair_all; do python -c "
import json
d=json.loads(open('gpurun_out/r02i_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['e2e']['ms_per_step'],3))"; done
python tools/train_step_probe.py > gpurun_out/r02i_train_probe.log 2>&1
python tools/gpu_debug.py prof 65536 > gpurun_out/r02i_prof_pair.log 2>&1
NFB_NO_PAIR=1 python tools/gpu_debug.py prof 65536 > gpurun_out/r02i_prof_nopair.log 2>&1
python -m pytest tests -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02i_pytest.log
tail -22 gpurun_out/r02i_pytest.log
grep step gpurun_out/r02i_train_probe.log
