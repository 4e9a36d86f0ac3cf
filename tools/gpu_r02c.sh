#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -x 2>&1 | tail -70 > gpurun_out/r02c_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-eager > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
NFB_NO_H2D_OVERLAP=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-eager > gpurun_out/r02c_bench_nooverlap.json 2>> gpurun_out/r02c_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02c_ref.json 2> gpurun_out/r02c_ref.err
tail -40 gpurun_out/r02c_pytest.log
