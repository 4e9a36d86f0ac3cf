#!/bin/bash
mkdir -p gpurun_out
T=r03d
( python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
python tools/bench_configs.py c3 2>&1 | tail -1
python tools/bench_configs.py c3 2>&1 | tail -1
