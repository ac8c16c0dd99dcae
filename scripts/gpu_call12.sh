#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call12.log
run2() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench_step.py --steps 30 "$@" 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3 | cut -c1-330; }
{
timeout 600 python -m pytest tests/test_gpu_corr.py tests/test_gpu_corr_fused.py tests/test_gpu_harness.py -q 2>&1 | tail -4
timeout 120 python scripts/microbench/time_tc.py fwd 2>&1 | tail -2
timeout 300 python bench_ops.py --quick --iters 20 2>&1 | grep -E "warp1d_bwd|fused f2|corr1d_fwd" | cut -c1-200
echo "--- 2 GPUs: peer exchange"; run2
echo "--- 2 GPUs: NCCL BN"; run2 --nccl-bn
echo "--- 2 GPUs: peer exchange, full depth"; run2 --full-depth
echo "--- 2 GPUs: NCCL BN, full depth"; run2 --full-depth --nccl-bn
} > $L 2>&1
cat $L
