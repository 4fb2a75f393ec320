#!/bin/bash
# Round 2, GPU call 37: lane refill for the shadow rays (trace_con_refill, busy-lane threshold 16 / 24 / 28) against the batch loop,
# on the scenes that run the refill kernels (from 1024 triangles on).
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
timeout 300 python tools/run_with_lib.py $V/libtracer_con24.so tools/gpu_parity_quick.py spectrumsphere spectrumspherehigh > $O/r2c37_parity_con24.log 2>&1; echo "con24 parity rc=$?"; grep -o '"[a-z0-9_]*": false' $O/r2c37_parity_con24.log | head -3
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py 3 4 5 > $O/r2c37_base_$rep.jsonl 2>/dev/null; echo base; cut -c1-130 $O/r2c37_base_$rep.jsonl
  for n in con16 con24 con28; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py 3 4 5 > $O/r2c37_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c37_${n}_$rep.jsonl
  done
done
