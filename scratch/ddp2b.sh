#!/bin/bash
mkdir -p gpurun_out
{
timeout 200 python -m pytest tests/test_gpu_harness.py -x -q 2>&1 | grep -v Warning | tail -15
N=$(nvidia-smi -L | wc -l); echo "gpus=$N"
for extra in "" "--no-pair"; do
echo "== $N GPUs graph $extra"
PMT_STEP_HANG_DUMP=110 timeout 140 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 bench_step.py --steps 30 $extra 2>&1 | grep -v "UserWarning\|run_backward\|^\*\|OMP_NUM" | tail -3
done
} > gpurun_out/ddp2b.log 2>&1
tail -30 gpurun_out/ddp2b.log
