#!/bin/bash
# round-2 first GPU call: full -m gpu suite (incl. the new 32-layer / trained-weight parity tests), the same two
# tests without the accumulate-truncation gain, bench (with reference-eager denominator), reference arm, pair MMA tool
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | tail -80 > gpurun_out/r02a_pytest.log
NFB_ACC_COMP_STEP=0 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "bench_config or trained" 2>&1 | tail -40 > gpurun_out/r02a_nogain.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02a_ref.json 2> gpurun_out/r02a_ref.err
timeout 120 ./tools/pair_mma_test > gpurun_out/r02a_pair.log 2>&1
nproc > gpurun_out/r02a_host.txt; lscpu | head -20 >> gpurun_out/r02a_host.txt
cat gpurun_out/r02a_pytest.log | tail -30
