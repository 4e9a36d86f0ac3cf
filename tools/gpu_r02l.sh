#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "gemm or backward" 2>&1 | tail -5 > gpurun_out/r02l_pytest.log
python tools/train_step_probe.py > gpurun_out/r02l_train_probe.log 2>&1
NFB_PROBE_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02l_train_launches.csv python tools/train_step_probe.py > gpurun_out/r02l_ncu_train.log 2>&1
tail -3 gpurun_out/r02l_pytest.log; grep step gpurun_out/r02l_train_probe.log
