#!/bin/bash
# Round 2, GPU call 14: evict-first path-state accesses, k_shade with a double-buffered exchange area (two barriers per iteration),
# per-launch statistics atomics, no emission lookup for dark materials.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c14_pytest_gpu.log 2>&1; tail -3 $O/r2c14_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c14_bench.json 2> $O/r2c14_bench.err; cut -c1-300 $O/r2c14_bench.json; tail -3 $O/r2c14_bench.err
timeout 600 python tools/bench_configs.py > $O/r2c14_configs.jsonl 2> $O/r2c14_configs.err; cut -c1-170 $O/r2c14_configs.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/r2c14_pass_launches.csv python tools/prof_pass.py cornell 1 > $O/r2c14_ncu_pass.log 2>&1
ls $O/r2c14_*
