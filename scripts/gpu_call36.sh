#!/bin/bash
# forward epilogue unrolled by three (register renaming instead of 32 copies per step): A/B + parity
mkdir -p gpurun_out
{
timeout 100 python scripts/microbench/ab_libs.py scratch/libpmt_ops_base.so pmt_learning_for_semantic_segmentation_and_disparity_b200/libpmt_ops.so 2>&1 | grep -v "^$" | cut -c1-250
echo "== pytest corr"; timeout 100 python -m pytest tests/test_gpu_corr.py tests/test_gpu_edge.py -q -m gpu --timeout 60 -x 2>&1 | tail -2
} > gpurun_out/r2_call36.log 2>&1
cat gpurun_out/r2_call36.log
