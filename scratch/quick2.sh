#!/bin/bash
mkdir -p gpurun_out
{
for i in 1 2; do
PMT_BWD_SPLIT=74 PMT_TC_DEBUG=2048 timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1
PMT_BWD_SPLIT=74 PMT_TC_DEBUG=4096 timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1
done
for s in 68 72 76; do PMT_BWD_SPLIT=$s timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1; done
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv
} > gpurun_out/quick2.log 2>&1
cat gpurun_out/quick2.log
