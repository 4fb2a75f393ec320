#!/bin/bash
# Round 2, GPU call 4 (first of the second session): the full -m gpu suite, bench, traversal variants on the large scenes
# (right-child prefetch, k_trace_sr, occupancy), ncu full captures of the traversal kernel on config 5.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c4_pytest_gpu.log 2>&1; tail -5 $O/r2c4_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c4_bench.json 2> $O/r2c4_bench.err; cut -c1-600 $O/r2c4_bench.json; tail -3 $O/r2c4_bench.err
for pf in 0 1 2; do
  LYS_TRACE_PF=$pf timeout 300 python tools/bench_configs.py 4 5 > $O/r2c4_configs_pf$pf.jsonl 2> $O/r2c4_configs_pf$pf.err
  cut -c1-170 $O/r2c4_configs_pf$pf.jsonl
done
for keep in 16 24; do
  LYS_TRACE_SR_KEEP=$keep LYS_TRACE_MODE=2 timeout 300 python tools/bench_configs.py 4 5 > $O/r2c4_configs_mode2_keep$keep.jsonl 2>/dev/null; cut -c1-170 $O/r2c4_configs_mode2_keep$keep.jsonl
done
LYS_TRACE_MODE=2 LYS_TRACE_SR_CAMERA=1 timeout 300 python tools/bench_configs.py 5 > $O/r2c4_configs_mode2_cam.jsonl 2>/dev/null; cut -c1-170 $O/r2c4_configs_mode2_cam.jsonl
for v in minb12 minb16 smem8; do
  timeout 300 python tools/run_with_lib.py $V/libtracer_$v.so tools/bench_configs.py 4 5 > $O/r2c4_configs_$v.jsonl 2> $O/r2c4_configs_$v.err
  cut -c1-170 $O/r2c4_configs_$v.jsonl
done
LYS_DETAIL=1 LYS_H=2160 LYS_W=3840 timeout 300 python tools/prof_pass.py synthetic 4 > $O/r2c4_synth_detail.log 2>&1; tail -4 $O/r2c4_synth_detail.log
LYS_TRACE_MODE=2 LYS_DETAIL=1 LYS_H=2160 LYS_W=3840 timeout 300 python tools/prof_pass.py synthetic 4 > $O/r2c4_synth_detail_mode2.log 2>&1; tail -4 $O/r2c4_synth_detail_mode2.log
LYS_H=2160 LYS_W=3840 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/r2c4_synth_pass_launches.csv python tools/prof_pass.py synthetic 1 > $O/r2c4_ncu_synth_pass.log 2>&1
LYS_H=2160 LYS_W=3840 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'^k_trace$' --launch-skip 16 --launch-count 2 -o $O/r2c4_synth_trace_full -f python tools/prof_pass.py synthetic 1 > $O/r2c4_ncu_synth_full.log 2>&1
tail -2 $O/r2c4_ncu_synth_full.log
ls -la $O/r2c4_*
