#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_harness.py -x -q 2>&1 | tail -8
echo "== 1 GPU graph paired"; timeout 200 python bench_step.py --steps 20 2>&1 | tail -1
echo "== 1 GPU graph unpaired"; timeout 200 python bench_step.py --steps 20 --no-pair 2>&1 | tail -1
} > gpurun_out/round9.log 2>&1
cat gpurun_out/round9.log
