#!/bin/bash
# Round 2, GPU call 12: L2 residency of the large BVH (persisting access-policy window on the pair records) and evict-first
# cache operators on the path-state records.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
timeout 900 python -m pytest tests -m gpu -x -q -k "million or baseline_resolution or full_size" > $O/r2c12_pytest_gpu.log 2>&1; tail -3 $O/r2c12_pytest_gpu.log
LYS_L2_WINDOW=0 timeout 300 python tools/bench_configs.py metric 4 5 > $O/r2c12_configs_win0.jsonl 2>/dev/null; cut -c1-170 $O/r2c12_configs_win0.jsonl
timeout 300 python tools/bench_configs.py metric 4 5 > $O/r2c12_configs_win1.jsonl 2>/dev/null; cut -c1-170 $O/r2c12_configs_win1.jsonl
LYS_L2_WINDOW=0 timeout 300 python tools/run_with_lib.py $V/libtracer_hints.so tools/bench_configs.py metric 3 4 5 > $O/r2c12_configs_hints_win0.jsonl 2>/dev/null; cut -c1-170 $O/r2c12_configs_hints_win0.jsonl
timeout 300 python tools/run_with_lib.py $V/libtracer_hints.so tools/bench_configs.py metric 3 4 5 > $O/r2c12_configs_hints_win1.jsonl 2>/dev/null; cut -c1-170 $O/r2c12_configs_hints_win1.jsonl
LYS_H=2160 LYS_W=3840 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file $O/r2c12_synth_pass_launches.csv python tools/prof_pass.py synthetic 1 > $O/r2c12_ncu_synth_pass.log 2>&1
ls $O/r2c12_*
