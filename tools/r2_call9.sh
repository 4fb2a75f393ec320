#!/bin/bash
# Round 2, GPU call 9: steps started ahead in a stepping loop -- parity suite, host-session fuzz on the device, interactive loop
# at depths 0..3 (same ARGB checksum required), breakdown, bench.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c9_pytest_gpu.log 2>&1; tail -5 $O/r2c9_pytest_gpu.log
timeout 300 python tools/fuzz_parity.py keys 11 150 > $O/r2c9_fuzz_keys.log 2>&1; tail -2 $O/r2c9_fuzz_keys.log
for d in 0 1 2 3; do
  LYS_STEP_AHEAD=$d timeout 300 python tools/bench_interactive.py cornell 1920 1080 600 > $O/r2c9_interactive_ahead$d.json 2>&1; tail -1 $O/r2c9_interactive_ahead$d.json
done
LYS_STEP_AHEAD=2 timeout 300 python tools/bench_interactive.py spectrumsphere 1920 1080 300 > $O/r2c9_interactive_sphere_ahead2.json 2>&1; tail -1 $O/r2c9_interactive_sphere_ahead2.json
LYS_STEP_AHEAD=0 timeout 300 python tools/bench_interactive.py spectrumsphere 1920 1080 300 > $O/r2c9_interactive_sphere_ahead0.json 2>&1; tail -1 $O/r2c9_interactive_sphere_ahead0.json
timeout 300 python tools/interactive_breakdown.py > $O/r2c9_interactive_breakdown.json 2>&1; tail -1 $O/r2c9_interactive_breakdown.json
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c9_bench.json 2> $O/r2c9_bench.err; cut -c1-300 $O/r2c9_bench.json; tail -3 $O/r2c9_bench.err
ls $O/r2c9_*
