#!/bin/bash
# round-2 evidence pass: gpu tests, bench line, ncu launch list, ncu --set full of the correlation kernels and of
# every other op, compute-sanitizer memcheck over the small-shape script
mkdir -p gpurun_out
L=gpurun_out/r2_call18.log
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -4
echo "== time_tc"; timeout 120 python scripts/microbench/time_tc.py fwd 2>&1 | tail -2; timeout 120 python scripts/microbench/time_tc.py bwd 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench.err | cut -c1-300
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench.json") if l.startswith("{")][-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "burst", "clocks")})
r = d["roofline"]; print({k: r[k] for k in ("achieved", "frac", "ms_per_launch", "kernel_sum_check")}); print(r["other_kernels"])
for o in d.get("ops", []): print(o)
print("step:", d.get("step"))
P
echo "== ncu launches"
PMT_BENCH_OPS=0 PMT_BENCH_STEP=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 5 --warmup 3 > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
echo "== ncu full corr"
PMT_BENCH_OPS=0 PMT_BENCH_STEP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'corr1d_(fwd|bwd)_tc' -s 8 -c 2 -o gpurun_out/r2_corr \
    python bench.py --steps 5 --warmup 3 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/r2_corr.ncu-rep gpurun_out/r2_ncu_corr.md > /dev/null 2>&1; echo "summary rc=$?"
echo "== ncu full ops"
timeout 300 python scripts/run_ops_once.py > gpurun_out/r2_ops_plain.log 2>&1 && tail -1 gpurun_out/r2_ops_plain.log &&
timeout 900 ncu --set full --clock-control none -k regex:'concat|softargmin|dispreg|upsample|warp_|bn_pair|corr_conv|corr2d' -c 80 -o /tmp/r2_ops python scripts/run_ops_once.py > gpurun_out/r2_ncu_ops.log 2>&1
tail -2 gpurun_out/r2_ncu_ops.log
ncu -i /tmp/r2_ops.ncu-rep --page raw --csv > gpurun_out/r2_ops_raw.csv 2>/dev/null; ls -la gpurun_out/r2_ops_raw.csv
echo "== sanitizer memcheck"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_small.py > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|all ops ran|Invalid|Error" gpurun_out/r2_memcheck.log | head -8
} > $L 2>&1
cat $L
