#!/bin/bash
# Round 2, GPU call 23: k_shade CTA size / occupancy without the 80-register cap (87-96 registers, no spills), A/B on one box.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
for n in s128b6 s128b5 s256b2 s192b4; do
  timeout 200 python tools/run_with_lib.py $V/libtracer_$n.so tools/gpu_parity_quick.py > $O/r2c23_parity_$n.log 2>&1; echo "$n parity rc=$?"; tail -1 $O/r2c23_parity_$n.log | cut -c1-200
done
for rep in 1 2 3; do
  timeout 300 python tools/bench_configs.py metric 2b 3 > $O/r2c23_base_$rep.jsonl 2>/dev/null; echo base; cut -c1-130 $O/r2c23_base_$rep.jsonl
  for n in s128b6 s128b5 s256b2 s192b4; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py metric 2b 3 > $O/r2c23_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c23_${n}_$rep.jsonl
  done
done
