#!/bin/bash
# Round 2, GPU call 38: shadow-ray lane refill on the large-scene layout only: busy-lane threshold 8 / 12 / 16 (build default) / 20 on
# config 5, parity (variant sweep with the large-scene layout forced on small scenes, 4K frame), launch list of the final pass.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
timeout 900 python -m pytest tests -m gpu -x -q -k "kernel_variants or baseline_resolution or million or soup or entry_points" > $O/r2c38_pytest_gpu.log 2>&1; tail -2 $O/r2c38_pytest_gpu.log
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py 5 > $O/r2c38_con16_$rep.jsonl 2>/dev/null; echo con16; cut -c1-130 $O/r2c38_con16_$rep.jsonl
  for n in con8 con12 con20; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py 5 > $O/r2c38_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c38_${n}_$rep.jsonl
  done
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio
LYS_H=2160 LYS_W=3840 timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/r2c38_synth_pass_launches.csv python tools/prof_pass.py synthetic 1 > $O/r2c38_ncu_synth_pass.log 2>&1
