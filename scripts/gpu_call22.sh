#!/bin/bash
# per-warp wait / section cycle counters of the first CTA of each gradient (profiling build), 3xTF32 and plain TF32
mkdir -p gpurun_out
L=gpurun_out/r2_call22.log
{
PMT_PROF_LIB=scratch/libpmt_ops_prof.so timeout 100 python scripts/microbench/prof_bwd.py 3
PMT_PROF_LIB=scratch/libpmt_ops_prof.so timeout 100 python scripts/microbench/prof_bwd.py 1
PMT_PROF_LIB=scratch/libpmt_ops_prof.so timeout 100 python scripts/microbench/trace_bwd.py
} > $L 2>&1
cat $L
