#!/bin/bash
PK=pmt_learning_for_semantic_segmentation_and_disparity_b200
mkdir -p gpurun_out
{
cp $PK/libpmt_ops.so /tmp/normal.so
cp scratch/libpmt_prof.so $PK/libpmt_ops.so
for d in 252 0; do PMT_BWD_SPLIT=74 PMT_TC_DEBUG=$d timeout 120 python scratch/trace_bwd.py; done
cp /tmp/normal.so $PK/libpmt_ops.so
} > gpurun_out/trace.log 2>&1
cat gpurun_out/trace.log
