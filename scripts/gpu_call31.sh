#!/bin/bash
# ablations of the final backward (development build): which removed piece of work shortens the kernel
mkdir -p gpurun_out
L=gpurun_out/r2_call31.log
D=scratch/libpmt_ops_dev.so
T="timeout 40 python scripts/microbench/time_bwd_modes.py $D"
{
$T "full"
PMT_TC_DEBUG=16 $T "no MMA"
PMT_TC_DEBUG=4 $T "no Gd build (zeros stored)"
PMT_TC_DEBUG=32 $T "no band split"
PMT_TC_DEBUG=8 $T "no epilogue"
PMT_TC_DEBUG=192 $T "no TMA traffic"
PMT_TC_DEBUG=60 $T "no MMA, build, split, epilogue"
PMT_PROF_LIB=scratch/libpmt_ops_prof.so timeout 40 python scripts/microbench/prof_bwd.py 3
} > $L 2>&1
cat $L
