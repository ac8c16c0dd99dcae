#!/bin/bash
# time the bwd TC kernel with different builder-warp counts (prebuilt .so variants)
for nb in 8 12 16; do
  cp scratch/libpmt_b$nb.so pmt_learning_for_semantic_segmentation_and_disparity_b200/libpmt_ops.so
  echo "builders=$nb"; timeout 100 python scratch/time_tc.py bwd 2>&1 | tail -2
  timeout 60 python scratch/test_tcb.py small 2>&1 | sed -n 3,4p
done
