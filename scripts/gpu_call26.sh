#!/bin/bash
# twice the band-split warps; tight timeouts, stop at the first failure
mkdir -p gpurun_out
L=gpurun_out/r2_call26.log
D=scratch/libpmt_ops_dev.so
T="timeout 40 python scripts/microbench/time_bwd_modes.py $D"
{
timeout 120 python scripts/microbench/ab_libs.py scratch/libpmt_ops_base.so pmt_learning_for_semantic_segmentation_and_disparity_b200/libpmt_ops.so 2>&1 | grep -v "^$" | grep "passes=3" | cut -c1-250
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "ab_libs failed or timed out"; exit 1; }
echo "== pytest corr"; timeout 150 python -m pytest tests/test_gpu_corr.py tests/test_gpu_edge.py tests/test_gpu_corr_fused.py -q -m gpu --timeout 60 -x 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "pytest failed or timed out"; exit 1; }
$T "both 74/74"
PMT_TC_DEBUG=2048 $T "gin1 only (74 SMs)"
PMT_TC_DEBUG=4096 $T "gin2 only (74 SMs)"
for s in 68 70 72 76; do PMT_BWD_SPLIT=$s $T "split $s/$((148-s))"; done
PMT_PROF_LIB=scratch/libpmt_ops_prof.so timeout 40 python scripts/microbench/prof_bwd.py 3
} > $L 2>&1
cat $L
