#!/bin/bash
# profiling build: where the MMA issuer's cycles go (sec0 = blocked in the MMA issue, sec1 = issue + commits + reconvergence)
mkdir -p gpurun_out
L=gpurun_out/r2_call33.log
{
PMT_PROF_LIB=scratch/libpmt_ops_prof.so timeout 40 python scripts/microbench/prof_bwd.py 3
PMT_TC_DEBUG=16 PMT_PROF_LIB=scratch/libpmt_ops_prof.so timeout 40 python scripts/microbench/prof_bwd.py 3
} > $L 2>&1
cat $L
