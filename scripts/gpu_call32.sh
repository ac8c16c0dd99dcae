#!/bin/bash
# look-ahead barrier tests in the MMA issuers (forward and backward): A/B against the committed round-1 build + parity tests
mkdir -p gpurun_out
L=gpurun_out/r2_call32.log
{
timeout 120 python scripts/microbench/ab_libs.py scratch/libpmt_ops_base.so pmt_learning_for_semantic_segmentation_and_disparity_b200/libpmt_ops.so 2>&1 | grep -v "^$" | cut -c1-250
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "ab_libs failed or timed out"; exit 1; }
echo "== pytest corr"; timeout 150 python -m pytest tests/test_gpu_corr.py tests/test_gpu_edge.py tests/test_gpu_corr_fused.py -q -m gpu --timeout 60 -x 2>&1 | tail -3
} > $L 2>&1
cat $L
