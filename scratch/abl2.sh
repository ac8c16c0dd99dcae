#!/bin/bash
set -u
PK=pmt_learning_for_semantic_segmentation_and_disparity_b200
mkdir -p gpurun_out
{
cp $PK/libpmt_ops.so /tmp/normal.so
cp scratch/libpmt_prof.so $PK/libpmt_ops.so
for d in 0 60; do PMT_TC_DEBUG=$d timeout 120 python scratch/prof_bwd.py 3; done
cp /tmp/normal.so $PK/libpmt_ops.so
} > gpurun_out/abl2.log 2>&1
cat gpurun_out/abl2.log
