#!/bin/bash
# ablations + per-role wait profile of the bwd TC kernel
set -u
PK=pmt_learning_for_semantic_segmentation_and_disparity_b200
mkdir -p gpurun_out
{
cp $PK/libpmt_ops.so /tmp/normal.so
cp scratch/libpmt_prof.so $PK/libpmt_ops.so
echo "== wait profile (3xTF32)"; timeout 120 python scratch/prof_bwd.py 3
cp /tmp/normal.so $PK/libpmt_ops.so
for d in 0 4 32 36 16 8 52 60 512; do
  PMT_TC_DEBUG=$d timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -2
done
for s in 60 64 68 72 74; do
  echo "split $s"; PMT_BWD_SPLIT=$s timeout 120 python scratch/time_tc.py bwd 2>&1 | tail -1
done
} > gpurun_out/abl.log 2>&1
cat gpurun_out/abl.log
