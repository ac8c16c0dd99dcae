#!/bin/bash
mkdir -p gpurun_out
{
for pat in 0 1 2 5 6; do for s in 2 4 8; do timeout 60 scripts/microbench/tma_bw $pat $s 256; done; done
for pat in 3 4; do for s in 2 4; do timeout 60 scripts/microbench/tma_bw $pat $s 256; done; done
echo "L2-resident (H=32)"
for pat in 0 1 2 3 4; do timeout 60 scripts/microbench/tma_bw $pat 4 32; done
echo "74 CTAs"
for pat in 0 1 2; do timeout 60 scripts/microbench/tma_bw $pat 4 256 74; done
} > gpurun_out/tma_bw.log 2>&1
cat gpurun_out/tma_bw.log
