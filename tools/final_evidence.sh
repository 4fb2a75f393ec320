#!/bin/bash
# Round evidence on ONE B200 (run through gpurun): tests, bench (both arms), all configs, per-bounce times, ncu launch
# list and full captures.  Everything lands in gpurun_out/; the summaries are copied to profiles/ by hand.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/final_pytest_gpu.log 2>&1; tail -2 $O/final_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $O/final_bench_n1.json 2> $O/final_bench_n1.err; cut -c1-200 $O/final_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_ref.json 2> $O/final_bench_ref.err; cut -c1-200 $O/final_bench_ref.json
LYS_DETAIL=1 python tools/prof_pass.py cornell 4 > $O/final_prof_pass.log 2>&1; tail -3 $O/final_prof_pass.log
python tools/bench_configs.py > $O/final_configs.jsonl 2> $O/final_configs.err; wc -l $O/final_configs.jsonl
# ncu: launch list of a short bench run (value printed under ncu is NOT a bench value)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/final_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/final_ncu_bench.log 2>&1
# ncu: launch list + DRAM bytes of one profiled pass (sequence of the shipped configuration, second pass so that the estimates exist)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/final_pass_launches.csv python tools/prof_pass.py cornell 1 > $O/final_ncu_pass.log 2>&1
# ncu: full captures of the steady-state pass (the first pass has 1 + 16 + 16 matching launches): generate+trace(-1), shade(0), trace(0), ..., k_tail
ncu --set full --import-source on --clock-control none -k regex:"k_generate_trace|k_trace|k_shade|k_tail" --launch-skip 33 --launch-count 12 -o $O/final_full -f python tools/prof_pass.py cornell 1 > $O/final_ncu_full.log 2>&1
ls -la $O/final_*
