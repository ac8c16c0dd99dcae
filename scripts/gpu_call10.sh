#!/bin/bash
mkdir -p gpurun_out profiles
L=gpurun_out/r2_call10.log
{
timeout 900 python -m pytest tests/test_gpu_corr.py -q -k "corr2d or general" 2>&1 | tail -8
timeout 300 python bench_ops.py --quick --iters 20 2>&1 | grep -E "warp|corr2d|fused f2" | cut -c1-190
timeout 300 python bench_step.py --steps 20 2>&1 | tail -1
timeout 300 python bench_step.py --steps 20 --unfused --no-lovasz 2>&1 | tail -1
timeout 300 python bench_step.py --steps 20 --full-depth 2>&1 | tail -1
timeout 120 python scripts/run_warp_bwd.py > gpurun_out/r2_warp_plain.log 2>&1 && cat gpurun_out/r2_warp_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:warp_bwd_rows -s 3 -c 1 -o gpurun_out/r2_warp_rows python scripts/run_warp_bwd.py > gpurun_out/r2_ncu_warp.log 2>&1
tail -3 gpurun_out/r2_ncu_warp.log
} > $L 2>&1
cat $L
