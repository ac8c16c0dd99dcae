#!/bin/bash
# gin2 with its A operand built straight into TMEM + band split teams; every step under a tight timeout, stop at the first failure
mkdir -p gpurun_out
L=gpurun_out/r2_call25.log
D=scratch/libpmt_ops_dev.so
T="timeout 40 python scripts/microbench/time_bwd_modes.py $D"
{
timeout 120 python scripts/microbench/ab_libs.py scratch/libpmt_ops_base.so pmt_learning_for_semantic_segmentation_and_disparity_b200/libpmt_ops.so 2>&1 | grep -v "^$" | tail -12 | cut -c1-250
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "ab_libs failed or timed out"; exit 1; }
echo "== pytest corr"; timeout 150 python -m pytest tests/test_gpu_corr.py tests/test_gpu_edge.py tests/test_gpu_corr_fused.py -q -m gpu --timeout 60 -x 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "pytest failed or timed out"; exit 1; }
$T "both 74/74"
PMT_TC_DEBUG=2048 $T "gin1 only (74 SMs)"
PMT_TC_DEBUG=4096 $T "gin2 only (74 SMs)"
PMT_BWD_TMEM_A1=0 $T "both, gin2 via smem A"
for s in 66 70 78 82; do PMT_BWD_SPLIT=$s $T "split $s/$((148-s))"; done
PASSES=1 $T "tf32 both"
PASSES=1 PMT_BWD_TMEM_A1=0 $T "tf32 both, gin2 via smem A"
PMT_PROF_LIB=scratch/libpmt_ops_prof.so timeout 40 python scripts/microbench/prof_bwd.py 3
} > $L 2>&1
cat $L
