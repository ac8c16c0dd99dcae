#!/bin/bash
# round-2 call 2: TMEM-A forward -- parity, timing vs the SS kernel (dev-knob rebuild on the box), aligned TMA row probe
mkdir -p gpurun_out
L=gpurun_out/r2_call2.log
{
timeout 600 python -m pytest tests/test_gpu_corr.py tests/test_gpu_harness.py -x -q 2>&1 | tail -4
timeout 120 python scripts/microbench/time_tc.py fwd 2>&1 | tail -2
timeout 120 scripts/microbench/tma_rows 2>&1
echo "--- dev-knob build: SS vs TMEM-A forward"
PMT_DEV_KNOBS=1 PMT_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for ss in 0 1; do echo "PMT_FWD_SS=$ss"; PMT_FWD_SS=$ss timeout 120 python scripts/microbench/time_tc.py fwd 2>&1 | tail -2; done
for lo in 1 2 3 4; do echo "TMEM-A lo_stages=$lo"; PMT_FWD_LO_STAGES=$lo timeout 120 python scripts/microbench/time_tc.py fwd 2>&1 | tail -1; done
} > $L 2>&1
cat $L
