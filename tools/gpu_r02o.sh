#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -k "glow or class_cond or conv2d or options" 2>&1 | tail -30 > gpurun_out/r02o_pytest.log
python tools/bench_configs.py c3 > gpurun_out/r02o_c3.json 2> gpurun_out/r02o_c3.err
tail -14 gpurun_out/r02o_pytest.log; cat gpurun_out/r02o_c3.json; tail -3 gpurun_out/r02o_c3.err
