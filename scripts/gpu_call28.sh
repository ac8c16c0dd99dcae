#!/bin/bash
# conflict-free rotated reads in the gin2 TMEM build: parity, A/B, SM split sweep (tight timeouts, stop at the first failure)
mkdir -p gpurun_out
L=gpurun_out/r2_call28.log
D=scratch/libpmt_ops_dev.so
T="timeout 40 python scripts/microbench/time_bwd_modes.py $D"
{
timeout 120 python scripts/microbench/ab_libs.py scratch/libpmt_ops_base.so pmt_learning_for_semantic_segmentation_and_disparity_b200/libpmt_ops.so 2>&1 | grep -v "^$" | cut -c1-250
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "ab_libs failed or timed out"; exit 1; }
echo "== pytest corr"; timeout 150 python -m pytest tests/test_gpu_corr.py tests/test_gpu_edge.py tests/test_gpu_corr_fused.py -q -m gpu --timeout 60 -x 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "pytest failed or timed out"; exit 1; }
$T "both default split"
PMT_TC_DEBUG=2048 PMT_BWD_SPLIT=74 $T "gin1 only (74 SMs)"
PMT_TC_DEBUG=4096 PMT_BWD_SPLIT=74 $T "gin2 only (74 SMs)"
for s in 70 74 76; do PMT_BWD_SPLIT=$s $T "split $s/$((148-s))"; done
PASSES=1 $T "tf32 both"
} > $L 2>&1
cat $L
