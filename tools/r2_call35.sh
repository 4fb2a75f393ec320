#!/bin/bash
# Round 2, GPU call 35: lane refill with octant copies on mid-size scenes (A/B), and the evidence of the refill kernels on config 5
# (all configs, ncu launch list and full capture of the 1 M-triangle pass).
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py 3 4 k21 k38 > $O/r2c35_base_$rep.jsonl 2>/dev/null; echo base; cut -c1-130 $O/r2c35_base_$rep.jsonl
  timeout 300 python tools/run_with_lib.py $V/libtracer_octrefill.so tools/bench_configs.py 3 4 k21 k38 > $O/r2c35_octrefill_$rep.jsonl 2>/dev/null; echo octrefill; cut -c1-130 $O/r2c35_octrefill_$rep.jsonl
done
timeout 600 python tools/bench_configs.py > $O/r2c35_configs.jsonl 2> $O/r2c35_configs.err; wc -l $O/r2c35_configs.jsonl
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio
LYS_H=2160 LYS_W=3840 timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/r2c35_synth_pass_launches.csv python tools/prof_pass.py synthetic 1 > $O/r2c35_ncu_synth_pass.log 2>&1
LYS_H=2160 LYS_W=3840 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'^k_trace$' --launch-skip 16 --launch-count 2 -o $O/r2c35_synth_trace_full -f python tools/prof_pass.py synthetic 1 > $O/r2c35_ncu_synth_full.log 2>&1
ls -la $O/r2c35_*
