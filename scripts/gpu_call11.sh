#!/bin/bash
# 2 GPUs: training step with the NVLink peer-memory BN exchange vs NCCL collectives; 1-GPU regression of the new kernels
mkdir -p gpurun_out
L=gpurun_out/r2_call11.log
run2() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench_step.py --steps 30 "$@" 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3; }
{
timeout 600 python -m pytest tests/test_gpu_warp_fused.py tests/test_gpu_corr_fused.py tests/test_gpu_harness.py -q 2>&1 | tail -4
timeout 300 python bench_ops.py --quick --iters 20 2>&1 | grep -E "warp1d_bwd|fused f2|blend" | cut -c1-200
echo "--- 2 GPUs: peer exchange"; run2
echo "--- 2 GPUs: NCCL BN"; run2 --nccl-bn
echo "--- 1 GPU"; timeout 300 python bench_step.py --steps 30 2>&1 | tail -1
} > $L 2>&1
cat $L
