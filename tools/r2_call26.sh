#!/bin/bash
# Round 2, GPU call 26: stackless single-box traversal (escape links, lbvh.cu k_thread_links): parity on the device, then A/B of
# prefetch levels / occupancy on top of it and of the single-box layout threshold (LYS_SINGLE_MAX) on configs 3 and 4.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c26_pytest_gpu.log 2>&1; tail -2 $O/r2c26_pytest_gpu.log
for rep in 1 2 3; do
  timeout 300 python tools/bench_configs.py metric 2b 3 4 5 > $O/r2c26_base_$rep.jsonl 2>/dev/null; echo base; cut -c1-130 $O/r2c26_base_$rep.jsonl
  for n in tpf2 tpf4 minb12; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py metric 2b 3 > $O/r2c26_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c26_${n}_$rep.jsonl
  done
  for sm in 4096 16384; do
    LYS_SINGLE_MAX=$sm timeout 300 python tools/bench_configs.py 3 4 > $O/r2c26_single${sm}_$rep.jsonl 2>/dev/null; echo single$sm; cut -c1-130 $O/r2c26_single${sm}_$rep.jsonl
  done
done
LYS_SINGLE_MAX=16384 timeout 200 python tools/gpu_parity_quick.py > $O/r2c26_parity_single16384.log 2>&1; echo "single16384 parity rc=$?"
