#!/bin/bash
# Round 2, GPU call 5: pair-node traversal -- parity suite, bench, configs, occupancy variants, NB by scene.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c5_pytest_gpu.log 2>&1; tail -5 $O/r2c5_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c5_bench.json 2> $O/r2c5_bench.err; cut -c1-300 $O/r2c5_bench.json; tail -3 $O/r2c5_bench.err
timeout 600 python tools/bench_configs.py > $O/r2c5_configs.jsonl 2> $O/r2c5_configs.err; cut -c1-170 $O/r2c5_configs.jsonl
LYS_TRACE_NB=1 timeout 300 python tools/bench_configs.py metric 2b 4 5 > $O/r2c5_configs_nb1.jsonl 2>/dev/null; cut -c1-170 $O/r2c5_configs_nb1.jsonl
LYS_TRACE_NB=2 timeout 300 python tools/bench_configs.py 3 4 5 > $O/r2c5_configs_nb2.jsonl 2>/dev/null; cut -c1-170 $O/r2c5_configs_nb2.jsonl
for v in minb8 minb12 minb16; do
  timeout 300 python tools/run_with_lib.py $V/libtracer_$v.so tools/bench_configs.py metric 4 5 > $O/r2c5_configs_$v.jsonl 2> $O/r2c5_configs_$v.err
  cut -c1-170 $O/r2c5_configs_$v.jsonl
done
LYS_DETAIL=1 LYS_H=2160 LYS_W=3840 timeout 300 python tools/prof_pass.py synthetic 4 > $O/r2c5_synth_detail.log 2>&1; tail -4 $O/r2c5_synth_detail.log
LYS_H=2160 LYS_W=3840 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/r2c5_synth_pass_launches.csv python tools/prof_pass.py synthetic 1 > $O/r2c5_ncu_synth_pass.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/r2c5_pass_launches.csv python tools/prof_pass.py cornell 1 > $O/r2c5_ncu_pass.log 2>&1
ls -la $O/r2c5_*
