#!/bin/bash
# training-step harness on every GPU of the box (graph-captured DDP + paired SyncBN); pass extra bench_step.py flags
mkdir -p gpurun_out
{
N=$(nvidia-smi -L | wc -l); echo "gpus=$N"
PMT_STEP_HANG_DUMP=110 timeout 140 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench_step.py --steps 30 "$@" 2>&1 | grep -v "UserWarning\|run_backward\|^\*\|OMP_NUM" | tail -5
} > gpurun_out/step_scaling.log 2>&1
tail -12 gpurun_out/step_scaling.log
