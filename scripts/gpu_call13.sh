#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call13.log
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench_step.py --steps 30 "$@" 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print({k: d[k] for k in ('n_gpus','ms_per_step','value','bn_exchange','full_depth','sync_bn','loss')})
    except Exception: print(l[:300])
"; }
{
timeout 300 python bench_ops.py --quick --iters 20 2>&1 | grep -E "warp|fused f2|production|config1|corr2d" | cut -c1-170
echo "--- 1 GPU plain (no process group)"; timeout 300 python bench_step.py --steps 30 --plain-single 2>&1 | tail -1 | cut -c1-200
echo "--- 1 GPU, 1-rank group, peer"; run 1
echo "--- 1 GPU, 1-rank group, nccl bn"; run 1 --nccl-bn
echo "--- 2 GPUs peer"; run 2
echo "--- 2 GPUs nccl"; run 2 --nccl-bn
echo "--- 2 GPUs no sync bn (diagnostic: DDP only)"; run 2 --no-sync-bn
echo "--- 1 GPU no sync bn"; run 1 --no-sync-bn
} > $L 2>&1
cat $L
