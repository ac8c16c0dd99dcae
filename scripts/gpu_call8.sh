#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call8.log
{
timeout 300 python scripts/debug_warp.py 2>&1 | tail -40
timeout 900 python -m pytest tests/test_gpu_warp.py tests/test_gpu_warp_fused.py tests/test_gpu_corr_fused.py -q 2>&1 | tail -25
timeout 300 python bench_ops.py --quick --iters 20 2>&1 | grep -E "warp" | cut -c1-170
} > $L 2>&1
cat $L
