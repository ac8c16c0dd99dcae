#!/bin/bash
# round-2 call 3: second-generation backward (A operand in TMEM for both gradients): parity, timing
mkdir -p gpurun_out
L=gpurun_out/r2_call3.log
{
timeout 300 python scripts/microbench/parity_tcb.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_corr.py -x -q 2>&1 | tail -15
timeout 120 python scripts/microbench/time_tc.py bwd 2>&1 | tail -2
timeout 300 python bench_ops.py --quick --iters 20 2>&1 | grep corr1d | cut -c1-260
} > $L 2>&1
cat $L
