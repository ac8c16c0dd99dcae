#!/bin/bash
# round-2 call 6: branch-free builders in the gen-2 backward; deterministic row-CSR warp backward
mkdir -p gpurun_out
L=gpurun_out/r2_call6.log
{
timeout 300 python scripts/microbench/parity_tcb.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_corr.py tests/test_gpu_warp.py tests/test_gpu_warp_fused.py tests/test_gpu_edge.py -x -q 2>&1 | tail -6
timeout 120 python scripts/microbench/time_tc.py fwd 2>&1 | tail -2
timeout 120 python scripts/microbench/time_tc.py bwd 2>&1 | tail -2
timeout 200 python bench_ops.py --quick --iters 20 2>&1 | grep -E "warp|corr1d" | cut -c1-200
echo "--- profile build"
PMT_BWD_PROFILE=1 PMT_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 120 python scripts/microbench/prof_tca.py 3 2>&1 | tail -20
for sp in 60 68 74; do echo "PMT_BWD_SPLIT=$sp"; PMT_BWD_SPLIT=$sp timeout 120 python scripts/microbench/time_tc.py bwd 2>&1 | tail -1; done
} > $L 2>&1
cat $L
