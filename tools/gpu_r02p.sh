#!/bin/bash
# baseline after container re-creation: full GPU suite + default bench
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/r02p_pytest.log 2>&1
( time python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err ) 2> gpurun_out/r02p_bench.time
tail -8 gpurun_out/r02p_pytest.log; cat gpurun_out/r02p_bench.json; tail -3 gpurun_out/r02p_bench.err; cat gpurun_out/r02p_bench.time
