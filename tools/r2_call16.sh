#!/bin/bash
# Round 2, GPU call 16: 256-bit loads of the traversal records; L1 carve-out experiment.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c16_pytest_gpu.log 2>&1; tail -3 $O/r2c16_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c16_bench.json 2> $O/r2c16_bench.err; cut -c1-300 $O/r2c16_bench.json; tail -3 $O/r2c16_bench.err
timeout 600 python tools/bench_configs.py > $O/r2c16_configs.jsonl 2> $O/r2c16_configs.err; cut -c1-170 $O/r2c16_configs.jsonl
LYS_L1_CARVEOUT=1 timeout 300 python tools/bench_configs.py 4 5 > $O/r2c16_configs_l1.jsonl 2>/dev/null; cut -c1-170 $O/r2c16_configs_l1.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/r2c16_pass_launches.csv python tools/prof_pass.py cornell 1 > $O/r2c16_ncu_pass.log 2>&1
ls $O/r2c16_*
