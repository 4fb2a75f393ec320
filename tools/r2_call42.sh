#!/bin/bash
# Round 2, GPU call 42: node stages per iteration of the lane-refill loops (1 / 2 = build default / 3 / 4).
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py 3 5 > $O/r2c42_rnb2_$rep.jsonl 2>/dev/null; echo rnb2; cut -c1-130 $O/r2c42_rnb2_$rep.jsonl
  for n in rnb1 rnb3 rnb4; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py 3 5 > $O/r2c42_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c42_${n}_$rep.jsonl
  done
done
