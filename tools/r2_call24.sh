#!/bin/bash
# Round 2, GPU call 24: software prefetch of the next batch's path-state records in k_shade / k_trace (CCTL.PF1), A/B on one box.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
for rep in 1 2 3; do
  timeout 300 python tools/bench_configs.py metric 2b 3 5 > $O/r2c24_base_$rep.jsonl 2>/dev/null; echo base; cut -c1-130 $O/r2c24_base_$rep.jsonl
  for n in pf tpf bothpf; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py metric 2b 3 5 > $O/r2c24_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c24_${n}_$rep.jsonl
  done
done
timeout 200 python tools/run_with_lib.py $V/libtracer_bothpf.so tools/gpu_parity_quick.py > $O/r2c24_parity_bothpf.log 2>&1; echo "bothpf parity rc=$?"
