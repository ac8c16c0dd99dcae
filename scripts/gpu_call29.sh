#!/bin/bash
# N-GPU check of the contract path after the child-group fix: bench.py under torchrun (headline + step record with the
# NVLink peer BN exchange); every step under a tight timeout
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
L=gpurun_out/r2_call29_n$N.log
{
echo "gpus=$N"
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n$N.json") if l.startswith("{")][-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "burst", "e2e")})
print("step:", d.get("step"))
P
tail -3 gpurun_out/step_n${N}_rank0.log | cut -c1-400
} > $L 2>&1
cat $L
