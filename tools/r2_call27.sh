#!/bin/bash
# Round 2, GPU call 27: stackless single-box records on LARGE scenes (one copy, select-based box test, LAY_SINGLE_SEL) against the
# pair records, at 16 / 12 / 10 resident CTAs per SM; and the octant-copy single-box layout on 20 K / 64 K-triangle scenes.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
LYS_OCT_ONE_COPY=1 LYS_BIG_SINGLE=1 timeout 200 python tools/gpu_parity_quick.py cornell spectrumsphere > $O/r2c27_parity_sel.log 2>&1; echo "sel parity rc=$?"
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py 5 > $O/r2c27_pair_$rep.jsonl 2>/dev/null; echo pair; cut -c1-130 $O/r2c27_pair_$rep.jsonl
  LYS_BIG_SINGLE=1 timeout 300 python tools/bench_configs.py 5 > $O/r2c27_sel16_$rep.jsonl 2>/dev/null; echo sel16; cut -c1-130 $O/r2c27_sel16_$rep.jsonl
  for n in sel12 sel10; do
    LYS_BIG_SINGLE=1 timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py 5 > $O/r2c27_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c27_${n}_$rep.jsonl
  done
  timeout 300 python tools/bench_configs.py k21 k38 > $O/r2c27_midpair_$rep.jsonl 2>/dev/null; echo midpair; cut -c1-130 $O/r2c27_midpair_$rep.jsonl
  LYS_SINGLE_MAX=100000 timeout 300 python tools/bench_configs.py k21 k38 > $O/r2c27_midsingle_$rep.jsonl 2>/dev/null; echo midsingle; cut -c1-130 $O/r2c27_midsingle_$rep.jsonl
done
