#!/bin/bash
# round-2 call 1: regression tests, new bench.py, TMA row probe, bounded attempt to install the upstream sampler
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -3 gpurun_out/r2_pytest1.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench1.json
timeout 120 scripts/microbench/tma_rows > gpurun_out/r2_tma_rows.log 2>&1; echo "tma_rows rc=$?"; cat gpurun_out/r2_tma_rows.log
( timeout 40 python -m pip download --no-deps -d /tmp/scs spatial-correlation-sampler 2>&1 | tail -5 ) > gpurun_out/r2_pip_attempt.log 2>&1; echo "pip rc=$?" >> gpurun_out/r2_pip_attempt.log
python -c "import spatial_correlation_sampler" >> gpurun_out/r2_pip_attempt.log 2>&1; cat gpurun_out/r2_pip_attempt.log
