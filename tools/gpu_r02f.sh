#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/pair_mma_test > gpurun_out/r02f_pair.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1
if ! grep -q "native backward ok" gpurun_out/r02f_smoke.log; then echo "PAIR SMOKE FAILED -> NFB_NO_PAIR=1" >> gpurun_out/r02f_smoke.log; export NFB_NO_PAIR=1; fi
python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r02f_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-eager --no-train-step > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
NFB_NO_PAIR=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-eager --no-train-step > gpurun_out/r02f_bench_nopair.json 2>> gpurun_out/r02f_bench.err
python tools/train_step_probe.py > gpurun_out/r02f_train_probe.log 2>&1
NFB_PROBE_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_train_launches.csv python tools/train_step_probe.py > gpurun_out/r02f_ncu_train.log 2>&1
cat gpurun_out/r02f_pair.log gpurun_out/r02f_smoke.log; tail -25 gpurun_out/r02f_pytest.log; cat gpurun_out/r02f_train_probe.log
