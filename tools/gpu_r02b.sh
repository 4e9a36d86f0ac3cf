#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r02b_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_run.py > gpurun_out/r02b_racecheck.log 2>&1
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_run.py > gpurun_out/r02b_synccheck.log 2>&1
tail -30 gpurun_out/r02b_pytest.log; tail -5 gpurun_out/r02b_racecheck.log; tail -5 gpurun_out/r02b_synccheck.log
cat /sys/fs/cgroup/cpu.max > gpurun_out/r02b_cpu.txt 2>&1; nproc >> gpurun_out/r02b_cpu.txt
for t in 2 4 16 32; do NFB_REF_THREADS=$t python bench.py --impl reference --steps 1 --warmup 1 --batch 16384 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threads', $t, d['value'], d['ms_per_step'], d['cpu_baseline']['sample'][:200])" >> gpurun_out/r02b_cpu.txt; done
cat gpurun_out/r02b_cpu.txt
