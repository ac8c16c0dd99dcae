#!/bin/bash
# full GPU round: smoke, -m gpu tests, bench, ncu launch list + full capture, per-op bench
bash scripts/gpu_round.sh ncu
timeout 300 python bench_ops.py > gpurun_out/ops.jsonl 2> gpurun_out/ops.err; echo "ops rc=$?"
