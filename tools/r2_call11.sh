#!/bin/bash
# Round 2, GPU call 11: phased traversal (budgeted phases + dense resume) on the pair layouts.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "kernel_variants or baseline_resolution" > $O/r2c11_pytest_gpu.log 2>&1; tail -5 $O/r2c11_pytest_gpu.log
for ph in 0 16 24 8,16 12,24 16,32 8,16,32 12,24,48; do
  LYS_TRACE_PHASES=$ph timeout 300 python tools/bench_configs.py 3 4 5 > $O/r2c11_configs_ph$ph.jsonl 2> $O/r2c11_configs_ph$ph.err; echo "phases $ph"; cut -c1-170 $O/r2c11_configs_ph$ph.jsonl; tail -2 $O/r2c11_configs_ph$ph.err
done
LYS_TRACE_PHASES=8,16,32 LYS_DETAIL=1 LYS_H=2160 LYS_W=3840 timeout 300 python tools/prof_pass.py synthetic 4 > $O/r2c11_synth_detail.log 2>&1; tail -4 $O/r2c11_synth_detail.log
LYS_TRACE_PHASES=8,16,32 LYS_H=2160 LYS_W=3840 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/r2c11_synth_pass_launches.csv python tools/prof_pass.py synthetic 1 > $O/r2c11_ncu_synth_pass.log 2>&1
ls $O/r2c11_*
