#!/bin/bash
# Round 2, GPU call 44: ncu launch list of the 1 M-triangle pass with the final kernels (refill loops with four node stages).
set -x
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio
LYS_H=2160 LYS_W=3840 timeout 300 ncu --metrics $M --clock-control none --csv --log-file $O/r2c44_synth_pass_launches.csv python tools/prof_pass.py synthetic 1 > $O/r2c44_ncu_synth_pass.log 2>&1
tail -1 $O/r2c44_ncu_synth_pass.log
