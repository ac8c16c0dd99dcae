#!/bin/bash
PK=pmt_learning_for_semantic_segmentation_and_disparity_b200
mkdir -p gpurun_out
{
cp $PK/libpmt_ops.so /tmp/normal.so
for nb in 8 12; do
  cp scratch/libpmt_b$nb.so $PK/libpmt_ops.so
  echo "builders=$nb"
  timeout 60 python scratch/test_tcb.py small 2>&1 | tail -3
  for s in 68 72 76; do PMT_BWD_SPLIT=$s timeout 40 python scratch/time_tc.py bwd 2>&1 | tail -1; done
  PMT_BWD_SPLIT=74 PMT_TC_DEBUG=2048 timeout 40 python scratch/time_tc.py bwd 2>&1 | tail -1
  PMT_BWD_SPLIT=74 PMT_TC_DEBUG=4096 timeout 40 python scratch/time_tc.py bwd 2>&1 | tail -1
done
cp /tmp/normal.so $PK/libpmt_ops.so
} > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
