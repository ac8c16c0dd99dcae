#!/bin/bash
# round-2 call 4: per-role wait counters of the second-generation backward (library built with PMT_BWD_PROFILE=1 here)
mkdir -p gpurun_out
L=gpurun_out/r2_call4.log
{
timeout 120 python scripts/microbench/prof_tca.py 3 2>&1 | tail -20
timeout 120 python scripts/microbench/prof_tca.py 1 2>&1 | tail -20
for sp in 60 74 88; do echo "PMT_BWD_SPLIT=$sp"; PMT_BWD_SPLIT=$sp timeout 120 python scripts/microbench/time_tc.py bwd 2>&1 | tail -1; done
} > $L 2>&1
cat $L
