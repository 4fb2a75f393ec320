#!/bin/bash
# Round 2, GPU call 41: shadow-ray refill with the finish records prefetched at acquisition (cpf), and the same on every refill scene (cpfall).
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py 3 4 5 > $O/r2c41_base_$rep.jsonl 2>/dev/null; echo base; cut -c1-130 $O/r2c41_base_$rep.jsonl
  timeout 300 python tools/run_with_lib.py $V/libtracer_cpf.so tools/bench_configs.py 5 > $O/r2c41_cpf_$rep.jsonl 2>/dev/null; echo cpf; cut -c1-130 $O/r2c41_cpf_$rep.jsonl
  timeout 300 python tools/run_with_lib.py $V/libtracer_cpfall.so tools/bench_configs.py 3 4 > $O/r2c41_cpfall_$rep.jsonl 2>/dev/null; echo cpfall; cut -c1-130 $O/r2c41_cpfall_$rep.jsonl
done
