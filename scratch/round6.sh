#!/bin/bash
mkdir -p gpurun_out
{
timeout 400 python -m pytest tests/test_gpu_corr.py -x -q 2>&1 | tail -2
for s in 70 74 78; do PMT_BWD_SPLIT=$s timeout 60 python scratch/time_tc.py bwd 2>&1 | tail -2; done
echo "passes=1 with 3 groups"
for s in 70 74; do PMT_BWD_GROUPS=3 PMT_BWD_SPLIT=$s timeout 60 python scratch/time_tc.py bwd 2>&1 | tail -2 | head -1; done
} > gpurun_out/round6.log 2>&1
cat gpurun_out/round6.log
