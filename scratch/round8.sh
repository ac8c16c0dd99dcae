#!/bin/bash
mkdir -p gpurun_out
{
for nb in 2 3; do echo "nbuf $nb"; PMT_FWD_NBUF=$nb timeout 60 python scratch/time_tc.py fwd 2>&1 | tail -2; done
} > gpurun_out/round8.log 2>&1
cat gpurun_out/round8.log
