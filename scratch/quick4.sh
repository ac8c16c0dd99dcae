#!/bin/bash
mkdir -p gpurun_out
{
echo "mode0 floor: all / no raw / no band / neither ; then full with no raw / no band / neither"
for d in 2108 2172 2236 2300 2112 2176 2240; do PMT_BWD_SPLIT=74 PMT_TC_DEBUG=$d timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1; done
echo "mode1 floor: all / no raw / no band / neither ; full: no raw / no band / neither"
for d in 4156 4220 4284 4348 4160 4224 4288; do PMT_BWD_SPLIT=74 PMT_TC_DEBUG=$d timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1; done
} > gpurun_out/quick4.log 2>&1
cat gpurun_out/quick4.log
