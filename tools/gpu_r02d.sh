#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -k "backward or autoregressive_sampling or reverse_kld" 2>&1 | tail -90 > gpurun_out/r02d_pytest.log
tail -60 gpurun_out/r02d_pytest.log
