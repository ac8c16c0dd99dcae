#!/bin/bash
mkdir -p gpurun_out
{
echo "mode0 full / floor / floor noprefetch / full noprefetch"
for d in 2048 2108 2620 2560; do PMT_BWD_SPLIT=74 PMT_TC_DEBUG=$d timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1; done
echo "mode0 NO_TMEM_A full / floor"
for d in 2048 2108; do PMT_NO_TMEM_A=1 PMT_BWD_SPLIT=74 PMT_TC_DEBUG=$d timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1; done
echo "mode0 at 148 CTAs full/floor"
for d in 2048 2108; do PMT_BWD_SPLIT=147 PMT_TC_DEBUG=$d timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1; done
} > gpurun_out/quick3.log 2>&1
cat gpurun_out/quick3.log
