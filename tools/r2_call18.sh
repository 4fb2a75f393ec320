#!/bin/bash
# Round 2, GPU call 18: node stages per loop iteration (1 / 2 / 3 / 4) on every config, interleaved repeats.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py metric 2b 3 4 5 > $O/r2c18_nb2_$rep.jsonl 2>/dev/null; echo nb2; cut -c1-130 $O/r2c18_nb2_$rep.jsonl
  for v in nb1 nb3 nb4; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$v.so tools/bench_configs.py metric 2b 3 4 5 > $O/r2c18_${v}_$rep.jsonl 2>/dev/null; echo $v; cut -c1-130 $O/r2c18_${v}_$rep.jsonl
  done
done
