#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_corr.py -x -q 2>&1 | tail -1
timeout 100 python scratch/time_tc.py fwd 2>&1 | tail -2
for s in 72 76 80 84; do echo "split $s"; PMT_BWD_SPLIT=$s timeout 100 python scratch/time_tc.py bwd 2>&1 | tail -2; done
} > gpurun_out/round5.log 2>&1
cat gpurun_out/round5.log
