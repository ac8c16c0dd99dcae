#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for s in 56 60 64 68 72; do echo "split $s"; PMT_BWD_SPLIT=$s timeout 100 python scratch/time_tc.py bwd 2>&1 | tail -2; done
timeout 100 python scratch/time_tc.py fwd 2>&1 | tail -2
} > gpurun_out/round4.log 2>&1
cat gpurun_out/round4.log
