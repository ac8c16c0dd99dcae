#!/bin/bash
# which gradient limits which generation of the tensor-core backward: each gradient alone on its 74 SMs, both generations;
# then the parity tests of the reworked builder loop and an ncu capture of the second generation
mkdir -p gpurun_out
L=gpurun_out/r2_call21.log
D=scratch/libpmt_ops_dev.so
T="timeout 100 python scripts/microbench/time_bwd_modes.py $D"
{
$T "gen1 both"
PMT_TC_DEBUG=2048 $T "gen1 gin1 only (74 SMs)"
PMT_TC_DEBUG=4096 $T "gen1 gin2 only (74 SMs)"
PMT_BWD_GEN2=1 $T "gen2 both"
PMT_BWD_GEN2=1 PMT_TCA_ONLY=0 $T "gen2 gin1 only (74 SMs)"
PMT_BWD_GEN2=1 PMT_TCA_ONLY=1 $T "gen2 gin2 only (74 SMs)"
PMT_BWD_SPLIT=64 $T "gen1 both split 64/84"
PMT_BWD_SPLIT=84 $T "gen1 both split 84/64"
PASSES=1 $T "gen1 both tf32"
PASSES=1 PMT_TC_DEBUG=2048 $T "gen1 gin1 only tf32"
PASSES=1 PMT_TC_DEBUG=4096 $T "gen1 gin2 only tf32"
echo "== pytest corr"; timeout 900 python -m pytest tests/test_gpu_corr.py tests/test_gpu_edge.py tests/test_gpu_corr_fused.py -q -m gpu --timeout 300 2>&1 | tail -4
echo "== ncu gen2"
PMT_BWD_GEN2=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr1d_bwd_tca -s 3 -c 1 -o gpurun_out/r2_gen2 python scripts/microbench/time_bwd_modes.py $D gen2 > gpurun_out/r2_ncu_gen2.log 2>&1; echo "ncu rc=$?"
} > $L 2>&1
cat $L
