#!/bin/bash
# N-GPU check of the contract path: bench.py under torchrun (headline + step record with the NVLink peer BN exchange),
# then the same step with one NCCL launch per BN layer for comparison
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
L=gpurun_out/r2_call19_n$N.log
{
echo "gpus=$N"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n$N.json") if l.startswith("{")][-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "burst", "e2e")})
print("step:", d.get("step"))
P
echo "--- $N GPUs nccl bn"
PMT_STEP_HANG_DUMP=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench_step.py --steps 30 --nccl-bn > gpurun_out/step_tmp.log 2>&1
grep -E '^\{' gpurun_out/step_tmp.log | tail -1 | tee gpurun_out/r2_step_ncclbn_n$N.json | cut -c1-700
} > $L 2>&1
cat $L
