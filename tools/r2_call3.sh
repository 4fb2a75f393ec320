#!/bin/bash
# Round 2, GPU call 3: occupancy / shared-stack variants on the large scenes, ncu full captures of the traversal kernels on
# config 5 (default and staged-refill), bench with the issue roofline.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
for v in minb12 minb16 smem8 smem16 smem24; do
  timeout 300 python tools/run_with_lib.py $V/libtracer_$v.so tools/bench_configs.py 4 5 > $O/r2c3_configs_$v.jsonl 2> $O/r2c3_configs_$v.err
  cut -c1-170 $O/r2c3_configs_$v.jsonl
done
LYS_TRACE_MODE=2 timeout 300 python tools/run_with_lib.py $V/libtracer_minb12.so tools/bench_configs.py 4 5 > $O/r2c3_configs_minb12_mode2.jsonl 2>/dev/null; cut -c1-170 $O/r2c3_configs_minb12_mode2.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c3_bench.json 2> $O/r2c3_bench.err; cut -c1-300 $O/r2c3_bench.json; tail -3 $O/r2c3_bench.err
LYS_H=2160 LYS_W=3840 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'^k_trace$' --launch-skip 16 --launch-count 2 -o $O/r2c3_synth_trace_full -f python tools/prof_pass.py synthetic 1 > $O/r2c3_ncu_synth_full.log 2>&1
tail -2 $O/r2c3_ncu_synth_full.log
LYS_TRACE_MODE=2 LYS_H=2160 LYS_W=3840 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'^k_trace_sr$' --launch-skip 16 --launch-count 2 -o $O/r2c3_synth_sr_full -f python tools/prof_pass.py synthetic 1 > $O/r2c3_ncu_synth_sr_full.log 2>&1
tail -2 $O/r2c3_ncu_synth_sr_full.log
ls -la $O/r2c3_*
