#!/bin/bash
# final round-2b evidence: full suite, full bench line, timeline, ncu launch list + full capture, sanitizers
mkdir -p gpurun_out
T=r03z
( time python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/${T}_pytest.log 2>&1
tail -5 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_full.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'train', d.get('train_step',{}).get('ms_per_step'))
print('ref eager', d.get('reference_eager_b200',{}).get('ms_per_step'), d.get('reference_eager_b200',{}).get('speedup_device'), d.get('reference_eager_b200',{}).get('speedup_e2e'), 'cpu', d.get('cpu_baseline',{}).get('value'))
for e in d.get('extra', []): print(e.get('config','')[:40], e.get('ms', e.get('ms_exact_2x2')))
PY
python tools/gpu_debug.py prof 65536 > gpurun_out/${T}_prof.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_rqs|diag_gauss|sum_stage|fill_kernel" -c 200 --csv --log-file gpurun_out/${T}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_rqs -s 3 -c 1 -o gpurun_out/${T}_fused_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-reference-eager --no-train-step --no-extra-configs > gpurun_out/${T}_ncu_full.log 2>&1
ls -la gpurun_out/${T}_fused_full.ncu-rep
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/${T}_racecheck.log 2>&1; tail -2 gpurun_out/${T}_racecheck.log
timeout 400 compute-sanitizer --tool synccheck python tools/sanitize_run.py > gpurun_out/${T}_synccheck.log 2>&1; tail -2 gpurun_out/${T}_synccheck.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
