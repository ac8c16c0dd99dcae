#!/bin/bash
# validate the pipelined forward epilogue + host row-block pipeline, time them, then ncu the non-headline kernels
mkdir -p gpurun_out
L=gpurun_out/r2_call16.log
{
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "--- time_tc fwd"; timeout 200 python scripts/microbench/time_tc.py fwd 2>&1 | tail -6
echo "--- bench_ops quick"; timeout 300 python bench_ops.py --quick 2>&1 | grep -E "warp|corr_conv|corr2d" | cut -c1-260
echo "--- bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -2 gpurun_out/r2_bench2.err; python - <<'P'
import json
d = json.loads(open("gpurun_out/r2_bench2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "burst", "roofline", "clocks")})
P
echo "--- ncu ops"
timeout 300 python scripts/run_ops_once.py > /dev/null 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:'concat|softargmin|dispreg|upsample|warp_|bn_pair|corr_conv|corr2d' -c 48 -o /tmp/r2_ops python scripts/run_ops_once.py > gpurun_out/r2_ncu_ops.log 2>&1
tail -2 gpurun_out/r2_ncu_ops.log
ncu -i /tmp/r2_ops.ncu-rep --page raw --csv > gpurun_out/r2_ops_raw.csv 2>/dev/null
ls -la /tmp/r2_ops.ncu-rep gpurun_out/r2_ops_raw.csv
} > $L 2>&1
cat $L
