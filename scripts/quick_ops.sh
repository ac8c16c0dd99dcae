#!/bin/bash
# quick GPU check of the streaming ops: parity tests + per-op bench
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_psmnet.py -x -q 2>&1 | tail -2
timeout 300 python bench_ops.py > gpurun_out/ops.jsonl 2> gpurun_out/ops.err; echo "ops rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/ops.jsonl'):
    d=json.loads(l)
    if any(k in d['op'] for k in ('concat','softargmin','dispreg','upsample')): print(d['op'],'|',d['config'][:24],'|',round(d['ms_per_launch']*1000,1),'us', round(d.get('hbm_frac') or 0,3))
P
} > gpurun_out/quick_ops.log 2>&1
cat gpurun_out/quick_ops.log
