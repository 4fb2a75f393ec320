#!/bin/bash
# Round 2, GPU call 22: k_shade with two barriers per iteration (alternating reflection queues) vs three, A/B on one box.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
timeout 600 python -m pytest tests -m gpu -x -q -k "pass_radiance or entry_points or soup or kernel_variants or baseline_resolution or stepping" > $O/r2c22_pytest_gpu.log 2>&1; tail -2 $O/r2c22_pytest_gpu.log
for rep in 1 2 3; do
  timeout 300 python tools/bench_configs.py metric 2b 3 5 > $O/r2c22_bar2_$rep.jsonl 2>/dev/null; echo bar2; cut -c1-130 $O/r2c22_bar2_$rep.jsonl
  timeout 300 python tools/run_with_lib.py $V/libtracer_bar3.so tools/bench_configs.py metric 2b 3 5 > $O/r2c22_bar3_$rep.jsonl 2>/dev/null; echo bar3; cut -c1-130 $O/r2c22_bar3_$rep.jsonl
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file $O/r2c22_pass_launches.csv python tools/prof_pass.py cornell 1 > $O/r2c22_ncu_pass.log 2>&1
