#!/bin/bash
# One GPU-box visit: smoke -> gpu tests -> bench -> ncu launch list -> ncu full capture of the corr kernels.
# Everything is wrapped in `timeout`; logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -5 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${1:-}" = "ncu" ]; then
  echo "== ncu launches"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
  echo "== ncu full"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:corr1d -s 8 -c 2 -o gpurun_out/prof \
      python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
  tail -3 gpurun_out/ncu_full.log
fi
