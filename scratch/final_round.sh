#!/bin/bash
bash scripts/gpu_round.sh ncu
timeout 300 python bench_ops.py > gpurun_out/ops.jsonl 2> gpurun_out/ops.err; echo "ops rc=$?"
