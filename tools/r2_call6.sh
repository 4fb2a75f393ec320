#!/bin/bash
# Round 2, GPU call 6: dual record layout (single-box records up to 1024 triangles, pair records above) -- parity suite, bench,
# all configs, NB / occupancy combinations on the pair layouts, interactive loop, ncu of the large scene.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c6_pytest_gpu.log 2>&1; tail -5 $O/r2c6_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c6_bench.json 2> $O/r2c6_bench.err; cut -c1-300 $O/r2c6_bench.json; tail -3 $O/r2c6_bench.err
timeout 600 python tools/bench_configs.py > $O/r2c6_configs.jsonl 2> $O/r2c6_configs.err; cut -c1-170 $O/r2c6_configs.jsonl
LYS_TRACE_NB=2 timeout 300 python tools/bench_configs.py 4 5 > $O/r2c6_configs_nb2.jsonl 2>/dev/null; cut -c1-170 $O/r2c6_configs_nb2.jsonl
LYS_TRACE_NB=1 timeout 300 python tools/bench_configs.py 3 > $O/r2c6_configs_nb1.jsonl 2>/dev/null; cut -c1-170 $O/r2c6_configs_nb1.jsonl
for v in minb12 minb16; do
  LYS_TRACE_NB=2 timeout 300 python tools/run_with_lib.py $V/libtracer_$v.so tools/bench_configs.py 3 4 5 > $O/r2c6_configs_${v}_nb2.jsonl 2> $O/r2c6_configs_$v.err
  cut -c1-170 $O/r2c6_configs_${v}_nb2.jsonl
done
timeout 300 python tools/bench_interactive.py > $O/r2c6_interactive.json 2>&1; tail -3 $O/r2c6_interactive.json
LYS_TRACE_NB=2 LYS_H=2160 LYS_W=3840 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/r2c6_synth_pass_launches.csv python tools/prof_pass.py synthetic 1 > $O/r2c6_ncu_synth_pass.log 2>&1
LYS_TRACE_NB=2 LYS_H=2160 LYS_W=3840 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'^k_trace$' --launch-skip 16 --launch-count 2 -o $O/r2c6_synth_trace_full -f python tools/prof_pass.py synthetic 1 > $O/r2c6_ncu_synth_full.log 2>&1
ls -la $O/r2c6_*
