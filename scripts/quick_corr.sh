#!/bin/bash
# quick GPU check of the correlation kernels: parity tests, then event timing at the headline shape
mkdir -p gpurun_out
{
timeout 400 python -m pytest tests/test_gpu_corr.py -x -q 2>&1 | tail -3
timeout 60 python scripts/microbench/time_tc.py fwd 2>&1 | tail -2
timeout 60 python scripts/microbench/time_tc.py bwd 2>&1 | tail -2
} > gpurun_out/quick_corr.log 2>&1
cat gpurun_out/quick_corr.log
