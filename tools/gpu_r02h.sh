#!/bin/bash
mkdir -p gpurun_out
NFB_NO_PAIR=1 python tools/gpu_debug.py prof 65536 > gpurun_out/r02h_prof_nopair.log 2>&1
python tools/gpu_debug.py prof 65536 > gpurun_out/r02h_prof_pair.log 2>&1
python tools/train_step_probe.py > gpurun_out/r02h_train_probe.log 2>&1
python -m pytest tests -m gpu -q -k "gemm or backward" 2>&1 | tail -5 > gpurun_out/r02h_pytest.log
cat gpurun_out/r02h_train_probe.log gpurun_out/r02h_pytest.log
