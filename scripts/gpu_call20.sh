#!/bin/bash
# A/B: committed library (scratch/libpmt_ops_base.so) vs the working tree (builder-loop rework in the backward,
# band-half-major MMA order in the forward), then the correlation parity tests on the new build
mkdir -p gpurun_out
L=gpurun_out/r2_call20.log
{
timeout 400 python scripts/microbench/ab_libs.py scratch/libpmt_ops_base.so pmt_learning_for_semantic_segmentation_and_disparity_b200/libpmt_ops.so 2>&1 | tail -20
echo "== pytest corr"; timeout 900 python -m pytest tests/test_gpu_corr.py tests/test_gpu_edge.py tests/test_gpu_corr_fused.py -q -m gpu --timeout 300 -x 2>&1 | tail -4
} > $L 2>&1
cat $L
