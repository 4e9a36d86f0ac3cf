#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_debug.py prof 65536 > gpurun_out/r02g_prof_pair.log 2>&1
python tools/gpu_debug.py prof 65536 > gpurun_out/r02g_prof_nopair.log 2>&1
python -m pytest tests -m gpu -q -s -k "options or neighbour or native_backward or gemm" 2>&1 | tail -40 > gpurun_out/r02g_pytest.log
tail -25 gpurun_out/r02g_pytest.log
