#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err ) 2> gpurun_out/r02m_bench.time
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02m_ref.json 2> gpurun_out/r02m_ref.err ) 2> gpurun_out/r02m_ref.time
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r02m_pytest.log
python tools/train_step_probe.py 2>/dev/null | grep step > gpurun_out/r02m_train_probe.log
cat gpurun_out/r02m_bench.time gpurun_out/r02m_ref.time gpurun_out/r02m_pytest.log gpurun_out/r02m_train_probe.log; tail -2 gpurun_out/r02m_bench.err
