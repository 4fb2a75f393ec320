#!/bin/bash
# Round 2, GPU call 15: A/B of the k_shade changes on one box (interleaved repeats).
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py metric 3 > $O/r2c15_base_$rep.jsonl 2>/dev/null; echo base; cut -c1-130 $O/r2c15_base_$rep.jsonl
  for v in sbuf nodark nohints sbuf_nodark; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$v.so tools/bench_configs.py metric 3 > $O/r2c15_${v}_$rep.jsonl 2>/dev/null; echo $v; cut -c1-130 $O/r2c15_${v}_$rep.jsonl
  done
done
