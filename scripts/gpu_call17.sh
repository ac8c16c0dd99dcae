#!/bin/bash
# step scaling at 8 GPUs (full-depth SDNetLite, graph-captured step): NVLink peer BN exchange vs one NCCL launch per layer
mkdir -p gpurun_out
L=gpurun_out/r2_call17.log
run() { n=$1; shift; PMT_STEP_HANG_DUMP=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench_step.py --steps 30 "$@" > gpurun_out/step_tmp.log 2>&1; grep -E '^\{' gpurun_out/step_tmp.log | tail -1 | tee -a gpurun_out/r2_step_scaling.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d.get(k) for k in ('n_gpus','ms_per_step','value','bn_exchange','full_depth','sync_bn','loss','nccl_launches_per_step_bn')})
"; grep -E "Error|error|Timeout|timed out" gpurun_out/step_tmp.log | head -3 | cut -c1-300; }
{
N=$(nvidia-smi -L | wc -l); echo "gpus=$N"
echo "--- $N GPUs peer"; run $N
echo "--- $N GPUs nccl bn"; run $N --nccl-bn
echo "--- 1 GPU peer (1-rank group)"; run 1
if [ "$N" -ge 8 ]; then echo "--- 4 GPUs peer"; run 4; fi
echo "--- bench.py under torchrun at $N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -2 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n$N.json") if l.startswith("{")][-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "step")})
P
} > $L 2>&1
cat $L
