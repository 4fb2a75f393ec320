#!/bin/bash
# Round 2, GPU call 10: overlapping ahead passes, per-leaf shading frame + single Schlick evaluation.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c10_pytest_gpu.log 2>&1; tail -5 $O/r2c10_pytest_gpu.log
timeout 300 python tools/fuzz_parity.py keys 12 150 > $O/r2c10_fuzz_keys.log 2>&1; tail -2 $O/r2c10_fuzz_keys.log
for d in 0 1 2 3; do
  LYS_STEP_AHEAD=$d timeout 300 python tools/bench_interactive.py cornell 1920 1080 600 > $O/r2c10_interactive_ahead$d.json 2>&1; tail -1 $O/r2c10_interactive_ahead$d.json
done
timeout 300 python tools/bench_interactive.py spectrumsphere 1920 1080 300 > $O/r2c10_interactive_sphere.json 2>&1; tail -1 $O/r2c10_interactive_sphere.json
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c10_bench.json 2> $O/r2c10_bench.err; cut -c1-300 $O/r2c10_bench.json; tail -3 $O/r2c10_bench.err
timeout 600 python tools/bench_configs.py > $O/r2c10_configs.jsonl 2> $O/r2c10_configs.err; cut -c1-170 $O/r2c10_configs.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/r2c10_pass_launches.csv python tools/prof_pass.py cornell 1 > $O/r2c10_ncu_pass.log 2>&1
ls $O/r2c10_*
