#!/bin/bash
# evidence: ncu --set full of every non-headline kernel at its BASELINE configuration
mkdir -p gpurun_out
L=gpurun_out/r2_call15.log
{
timeout 300 python scripts/run_ops_once.py > gpurun_out/r2_ops_plain.log 2>&1 && cat gpurun_out/r2_ops_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:'concat|softargmin|dispreg|upsample|warp_|bn_pair|corr_conv|corr2d' -c 60 -o gpurun_out/r2_ops python scripts/run_ops_once.py > gpurun_out/r2_ncu_ops.log 2>&1
tail -3 gpurun_out/r2_ncu_ops.log
} > $L 2>&1
cat $L
