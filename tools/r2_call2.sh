#!/bin/bash
# Round 2, GPU call 2: full -m gpu suite with the new BASELINE-size tests, bench with the new roofline, right-child prefetch
# variants on the large scenes, ncu launch list of the production Cornell pass, ncu full captures of k_trace on config 5.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c2_pytest_gpu.log 2>&1; tail -5 $O/r2c2_pytest_gpu.log
timeout 300 python tools/fuzz_parity.py lbvh 22 300 > $O/r2c2_fuzz_lbvh.log 2>&1; tail -2 $O/r2c2_fuzz_lbvh.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c2_bench.json 2> $O/r2c2_bench.err; cut -c1-400 $O/r2c2_bench.json; tail -3 $O/r2c2_bench.err
for pf in 0 1 2 3; do
  LYS_TRACE_PF=$pf timeout 300 python tools/bench_configs.py 3 4 5 > $O/r2c2_configs_pf$pf.jsonl 2> $O/r2c2_configs_pf$pf.err
  cut -c1-170 $O/r2c2_configs_pf$pf.jsonl
done
LYS_TRACE_PF=1 LYS_TRACE_MODE=2 timeout 300 python tools/bench_configs.py 5 > $O/r2c2_configs_pf1_mode2.jsonl 2>/dev/null; cut -c1-170 $O/r2c2_configs_pf1_mode2.jsonl
# ncu launch list of the production sequence (CornellBox 1080p, second pass = steady state)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/r2c2_pass_launches.csv python tools/prof_pass.py cornell 1 > $O/r2c2_ncu_pass.log 2>&1
# ncu full set: k_trace(0), k_trace(1) of the steady-state pass on config 5 (the warm-up pass has 16 k_trace launches)
LYS_H=2160 LYS_W=3840 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_trace<" --launch-skip 16 --launch-count 2 -o $O/r2c2_synth_trace_full -f python tools/prof_pass.py synthetic 1 > $O/r2c2_ncu_synth_full.log 2>&1
tail -2 $O/r2c2_ncu_synth_full.log
ls -la $O/r2c2_*
