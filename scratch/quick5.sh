#!/bin/bash
mkdir -p gpurun_out
{
for ns in 0 32 64 128 256; do echo "relax $ns"; PMT_BWD_RELAX_NS=$ns timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -2; done
echo "mode0 / mode1 alone relax 64 vs 0"
for ns in 0 64; do
PMT_BWD_RELAX_NS=$ns PMT_BWD_SPLIT=74 PMT_TC_DEBUG=2048 timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1
PMT_BWD_RELAX_NS=$ns PMT_BWD_SPLIT=74 PMT_TC_DEBUG=4096 timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1
done
} > gpurun_out/quick5.log 2>&1
cat gpurun_out/quick5.log
