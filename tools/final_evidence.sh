#!/bin/bash
# Round evidence on ONE B200 (run through gpurun): tests, device fuzz, bench (both arms), all configs, interactive loop,
# ncu launch lists (CornellBox pass, 1 M-triangle pass, 1 M-triangle build) and full captures of the dominant kernels.
# Everything lands in gpurun_out/final_*; the summaries are copied to profiles/ by hand.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/final_pytest_gpu.log 2>&1; tail -2 $O/final_pytest_gpu.log
timeout 300 python tools/fuzz_parity.py lbvh 21 200 > $O/final_fuzz.log 2>&1; timeout 300 python tools/fuzz_parity.py soup 22 100 >> $O/final_fuzz.log 2>&1; timeout 300 python tools/fuzz_parity.py keys 23 150 >> $O/final_fuzz.log 2>&1; grep -c . $O/final_fuzz.log; grep "scenes" $O/final_fuzz.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/final_bench_n1.json 2> $O/final_bench_n1.err; cut -c1-200 $O/final_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_ref.json 2> $O/final_bench_ref.err; cut -c1-200 $O/final_bench_ref.json
timeout 600 python tools/bench_configs.py > $O/final_configs.jsonl 2> $O/final_configs.err; wc -l $O/final_configs.jsonl
timeout 300 python tools/bench_interactive.py cornell 1920 1080 600 > $O/final_interactive.json 2>&1; tail -1 $O/final_interactive.json
LYS_DETAIL=1 timeout 300 python tools/prof_pass.py cornell 4 > $O/final_prof_pass.log 2>&1; tail -3 $O/final_prof_pass.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio
# launch list of a short bench run (a value printed under ncu is NOT a bench value)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/final_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/final_ncu_bench.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/final_pass_launches.csv python tools/prof_pass.py cornell 1 > $O/final_ncu_pass.log 2>&1
LYS_H=2160 LYS_W=3840 timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/final_synth_pass_launches.csv python tools/prof_pass.py synthetic 1 > $O/final_ncu_synth_pass.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/final_build_1M_launches.csv python tools/build_1m.py > $O/final_ncu_build.log 2>&1
# full captures of the steady-state CornellBox pass (the warm-up pass has 1 + 16 + 16 matching launches): generate+trace(-1), shade(0), trace(0), shade(1), trace(1)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_generate_trace|k_trace|k_shade|k_tail" --launch-skip 33 --launch-count 5 -o $O/final_full -f python tools/prof_pass.py cornell 1 > $O/final_ncu_full.log 2>&1
# and of trace(0), trace(1) on the 1 M-triangle scene
LYS_H=2160 LYS_W=3840 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'^k_trace$' --launch-skip 16 --launch-count 2 -o $O/final_synth_trace_full -f python tools/prof_pass.py synthetic 1 > $O/final_ncu_synth_full.log 2>&1
ls -la $O/final_*
