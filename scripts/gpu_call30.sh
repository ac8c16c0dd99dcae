#!/bin/bash
# validation of the reworked backward: all GPU tests, A/B timing, the bench line, ncu launch list and full capture
mkdir -p gpurun_out
L=gpurun_out/r2_call30.log
{
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 400 python -m pytest tests -q -m gpu --timeout 120 -x 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "pytest failed or timed out"; exit 1; }
timeout 120 python scripts/microbench/ab_libs.py scratch/libpmt_ops_base.so pmt_learning_for_semantic_segmentation_and_disparity_b200/libpmt_ops.so 2>&1 | grep -v "^$" | grep "passes=3" | cut -c1-250
echo "== bench"; timeout 500 python bench.py --steps 100 --warmup 5 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_c.err | cut -c1-300
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_c.json") if l.startswith("{")][-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "burst", "clocks")})
r = d["roofline"]; print({k: r[k] for k in ("achieved", "frac", "ms_per_launch", "kernel_sum_check")}); print(r["other_kernels"])
print("step:", d.get("step"))
P
echo "== ncu launches"
PMT_BENCH_OPS=0 PMT_BENCH_STEP=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_c.csv \
    python bench.py --steps 5 --warmup 3 > gpurun_out/r2_ncu_launch_c.log 2>&1; echo "ncu launches rc=$?"
echo "== ncu full corr"
PMT_BENCH_OPS=0 PMT_BENCH_STEP=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'corr1d_(fwd|bwd)_tc' -s 8 -c 2 -o gpurun_out/r2_corr_c \
    python bench.py --steps 5 --warmup 3 > gpurun_out/r2_ncu_full_c.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/r2_corr_c.ncu-rep gpurun_out/r2_ncu_corr_c.md > /dev/null 2>&1; echo "summary rc=$?"
} > $L 2>&1
cat $L
