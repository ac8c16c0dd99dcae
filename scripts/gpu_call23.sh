#!/bin/bash
# band split on dedicated warps: A/B against the committed build (bit-exact check + timing), parity tests, SM split sweep,
# per-warp cycle counters
mkdir -p gpurun_out
L=gpurun_out/r2_call23.log
D=scratch/libpmt_ops_dev.so
T="timeout 100 python scripts/microbench/time_bwd_modes.py $D"
{
timeout 300 python scripts/microbench/ab_libs.py scratch/libpmt_ops_base.so pmt_learning_for_semantic_segmentation_and_disparity_b200/libpmt_ops.so 2>&1 | grep -v "^$" | tail -12 | cut -c1-250
echo "== pytest corr"; timeout 900 python -m pytest tests/test_gpu_corr.py tests/test_gpu_edge.py tests/test_gpu_corr_fused.py -q -m gpu --timeout 300 2>&1 | tail -3
$T "both 74/74"
PMT_TC_DEBUG=2048 $T "gin1 only (74 SMs)"
PMT_TC_DEBUG=4096 $T "gin2 only (74 SMs)"
for s in 58 62 66 70 78; do PMT_BWD_SPLIT=$s $T "split $s/$((148-s))"; done
PMT_PROF_LIB=scratch/libpmt_ops_prof.so timeout 100 python scripts/microbench/prof_bwd.py 3
} > $L 2>&1
cat $L
