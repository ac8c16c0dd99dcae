#!/bin/bash
mkdir -p gpurun_out
{
for ls in 1 2 3 4; do echo "lo_stages $ls"; PMT_FWD_LO_STAGES=$ls timeout 60 python scratch/time_tc.py fwd 2>&1 | tail -1; done
} > gpurun_out/round8.log 2>&1
cat gpurun_out/round8.log
