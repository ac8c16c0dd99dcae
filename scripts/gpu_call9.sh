#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call9.log
{
timeout 300 python scripts/debug_warp.py 2>&1 | tail -14
timeout 900 python -m pytest tests/test_gpu_warp.py tests/test_gpu_warp_fused.py tests/test_gpu_corr_fused.py tests/test_gpu_harness.py -q 2>&1 | tail -25
timeout 300 python bench_ops.py --quick --iters 20 2>&1 | grep -E "warp" | cut -c1-170
timeout 300 python bench_step.py --steps 20 2>&1 | tail -2
timeout 300 python bench_step.py --steps 20 --unfused --no-lovasz 2>&1 | tail -1
timeout 300 python bench_step.py --steps 20 --full-depth 2>&1 | tail -1
} > $L 2>&1
cat $L
