#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call14.log
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench_step.py --steps 30 "$@" > gpurun_out/step_tmp.log 2>&1; grep -E '^\{' gpurun_out/step_tmp.log | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('n_gpus','ms_per_step','value','bn_exchange','full_depth','sync_bn','loss','nccl_launches_per_step_bn')})
"; grep -E "Error|error" gpurun_out/step_tmp.log | head -3 | cut -c1-300; }
{
timeout 600 python -m pytest tests/test_gpu_harness.py -q 2>&1 | tail -3
echo "--- 1 GPU, 1-rank group, peer"; run 1
echo "--- 1 GPU, 1-rank group, nccl bn"; run 1 --nccl-bn
echo "--- 2 GPUs peer"; run 2
echo "--- 2 GPUs nccl"; run 2 --nccl-bn
echo "--- 2 GPUs peer full depth"; run 2 --full-depth
echo "--- 2 GPUs nccl full depth"; run 2 --full-depth --nccl-bn
echo "--- 1 GPU peer full depth"; run 1 --full-depth
} > $L 2>&1
cat $L
