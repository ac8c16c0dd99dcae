#!/bin/bash
# round-2 call 5: elect-based MMA issue in all tensor-core kernels: parity, timing, then per-role counters (profile build on the box)
mkdir -p gpurun_out
L=gpurun_out/r2_call5.log
{
timeout 300 python scripts/microbench/parity_tcb.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_corr.py -x -q 2>&1 | tail -4
timeout 120 python scripts/microbench/time_tc.py fwd 2>&1 | tail -2
timeout 120 python scripts/microbench/time_tc.py bwd 2>&1 | tail -2
echo "--- profile build"
PMT_BWD_PROFILE=1 PMT_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 120 python scripts/microbench/prof_tca.py 3 2>&1 | tail -20
for ss in 0 1; do echo "PMT_FWD_SS=$ss"; PMT_FWD_SS=$ss timeout 120 python scripts/microbench/time_tc.py fwd 2>&1 | tail -2; done
for v1 in 0 1; do echo "PMT_BWD_V1=$v1"; PMT_BWD_V1=$v1 timeout 120 python scripts/microbench/time_tc.py bwd 2>&1 | tail -2; done
for sp in 64 68 80; do echo "PMT_BWD_SPLIT=$sp"; PMT_BWD_SPLIT=$sp timeout 120 python scripts/microbench/time_tc.py bwd 2>&1 | tail -1; done
} > $L 2>&1
cat $L
