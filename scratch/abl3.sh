#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "--- mode 0 only (2048), 74 CTAs"
for d in 2048 2052 2080 2064 2056 2084 2100 2108; do PMT_BWD_SPLIT=74 PMT_TC_DEBUG=$d timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1; done
echo "--- mode 1 only (4096), 74 CTAs"
for d in 4096 4100 4128 4112 4104 4132 4148 4156; do PMT_BWD_SPLIT=74 PMT_TC_DEBUG=$d timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1; done
} > gpurun_out/abl3.log 2>&1
cat gpurun_out/abl3.log
