#!/bin/bash
PK=pmt_learning_for_semantic_segmentation_and_disparity_b200
mkdir -p gpurun_out
{
cp $PK/libpmt_ops.so /tmp/normal.so
for rep in 1 2; do
for v in x1 x2; do
  cp scratch/libpmt_$v.so $PK/libpmt_ops.so
  echo "variant=$v"
  timeout 60 python scratch/time_tc.py fwd 2>&1 | tail -1
done
done
cp scratch/libpmt_x2.so $PK/libpmt_ops.so
timeout 300 python -m pytest tests/test_gpu_corr.py tests/test_gpu_edge.py tests/test_gpu_harness.py -x -q 2>&1 | tail -3
cp /tmp/normal.so $PK/libpmt_ops.so
} > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
