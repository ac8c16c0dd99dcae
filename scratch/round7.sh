#!/bin/bash
mkdir -p gpurun_out
{
timeout 400 python -m pytest tests/test_gpu_corr.py -x -q 2>&1 | tail -8
timeout 60 python scratch/time_tc.py fwd 2>&1 | tail -2
for d in 4 8 16 28; do PMT_TC_DEBUG=$d timeout 60 python scratch/time_tc.py fwd 2>&1 | tail -1; done
} > gpurun_out/round7.log 2>&1
cat gpurun_out/round7.log
