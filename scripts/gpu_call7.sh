#!/bin/bash
# round-2 call 7: parity of everything new (corr gens, warp rows + fused consumers, f2), timings old vs new backward
mkdir -p gpurun_out
L=gpurun_out/r2_call7.log
{
timeout 900 python -m pytest tests/test_gpu_corr.py tests/test_gpu_warp.py tests/test_gpu_warp_fused.py tests/test_gpu_corr_fused.py tests/test_gpu_edge.py -q 2>&1 | tail -25
timeout 300 python bench_ops.py --quick --iters 20 2>&1 | grep -E "warp|corr1d" | cut -c1-170
echo "--- dev-knob build"
PMT_DEV_KNOBS=1 PMT_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for v1 in 1 0; do echo "PMT_BWD_GEN2=$v1"; PMT_BWD_GEN2=$v1 timeout 120 python scripts/microbench/time_tc.py bwd 2>&1 | tail -2; done
PMT_BWD_GEN2=1 timeout 200 python scripts/microbench/parity_tcb.py 2>&1 | tail -12
} > $L 2>&1
cat $L
