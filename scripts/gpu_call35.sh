#!/bin/bash
# last check of the round: smoke + all GPU tests on the final build
mkdir -p gpurun_out
{
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python -m pytest tests -q -m gpu --timeout 100 -x 2>&1 | tail -2
} > gpurun_out/r2_call35.log 2>&1
cat gpurun_out/r2_call35.log
