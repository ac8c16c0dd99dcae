#!/bin/bash
# quick correctness + timing of the corr kernels
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_corr.py -x -q 2>&1 | tail -3
timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -2
for s in 60 64 68 72; do PMT_BWD_SPLIT=$s timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1; done
PMT_BWD_SPLIT=74 PMT_TC_DEBUG=2048 timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1
PMT_BWD_SPLIT=74 PMT_TC_DEBUG=4096 timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1
} > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
