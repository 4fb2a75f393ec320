#!/bin/bash
# Round 2, GPU call 33: lane refill for the closest-hit walk (stackless state), ext items only, hits-first order lists off in both arms.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
LYS_SHADE_ORDER=0 timeout 200 python tools/run_with_lib.py $V/libtracer_refill16.so tools/gpu_parity_quick.py cornell spectrumsphere > $O/r2c33_parity_refill16.log 2>&1; echo "refill16 parity rc=$?"; grep -o '"[a-z0-9_]*": false' $O/r2c33_parity_refill16.log | head -3
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py metric 3 5 > $O/r2c33_ordered_$rep.jsonl 2>/dev/null; echo ordered; cut -c1-130 $O/r2c33_ordered_$rep.jsonl
  LYS_SHADE_ORDER=0 timeout 300 python tools/bench_configs.py metric 3 5 > $O/r2c33_unordered_$rep.jsonl 2>/dev/null; echo unordered; cut -c1-130 $O/r2c33_unordered_$rep.jsonl
  for n in refill8 refill16 refill24; do
    LYS_SHADE_ORDER=0 timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py metric 3 5 > $O/r2c33_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c33_${n}_$rep.jsonl
  done
done
