#!/bin/bash
# Round 2, GPU call 28: node stages per loop iteration (TRAV_NB 1 / 2 / 3 / 4) of the stackless single-box walk, small and large scenes.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
export LYS_BIG_SINGLE=1
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py metric 3 5 > $O/r2c28_nb2_$rep.jsonl 2>/dev/null; echo nb2; cut -c1-130 $O/r2c28_nb2_$rep.jsonl
  for n in nb1 nb3 nb4; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py metric 3 5 > $O/r2c28_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c28_${n}_$rep.jsonl
  done
done
