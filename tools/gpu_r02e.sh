#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r02e_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-eager > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -40 gpurun_out/r02e_pytest.log; tail -3 gpurun_out/r02e_bench.err
