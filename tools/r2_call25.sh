#!/bin/bash
# Round 2, GPU call 25: k_trace prefetch levels (1: next batch; 2: + queue line of the current connect batch; 3: + early pid load and acc prefetch), A/B on one box.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
for rep in 1 2 3; do
  timeout 300 python tools/bench_configs.py metric 2b 3 5 > $O/r2c25_base_$rep.jsonl 2>/dev/null; echo base; cut -c1-130 $O/r2c25_base_$rep.jsonl
  for n in tpf1 tpf2 tpf3; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py metric 2b 3 5 > $O/r2c25_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c25_${n}_$rep.jsonl
  done
done
timeout 200 python tools/run_with_lib.py $V/libtracer_tpf3.so tools/gpu_parity_quick.py > $O/r2c25_parity_tpf3.log 2>&1; echo "tpf3 parity rc=$?"
