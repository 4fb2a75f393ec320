#!/bin/bash
# Round 2, GPU call 34: lane refill of the closest-hit walk with order lists (LAY_SEL): parity on the device (variant sweep, BASELINE
# resolutions incl. the 1 003 244-triangle scene at 4K), then the busy-lane threshold on config 5.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
timeout 900 python -m pytest tests -m gpu -x -q -k "kernel_variants or baseline_resolution or million or full_size or soup" > $O/r2c34_pytest_gpu.log 2>&1; tail -2 $O/r2c34_pytest_gpu.log
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py 5 > $O/r2c34_r24_$rep.jsonl 2>/dev/null; echo r24; cut -c1-130 $O/r2c34_r24_$rep.jsonl
  for n in r20 r28 r31; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py 5 > $O/r2c34_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c34_${n}_$rep.jsonl
  done
done
