#!/bin/bash
# Round 2, GPU call 29: warp-autonomous k_shade_w (LYS_SHADE_W=1: deferred reflection vertices, no block barrier) against the
# CTA-queue k_shade; resident CTAs per SM of the octant-copy traversal kernels on mid-size scenes.
set -x
O=gpurun_out
mkdir -p $O
V=msc-futhark-ray-tracer_b200/variants
LYS_SHADE_W=1 timeout 200 python tools/gpu_parity_quick.py > $O/r2c29_parity_shadew.log 2>&1; echo "shade_w parity rc=$?"; grep -c true $O/r2c29_parity_shadew.log; grep -o '"[a-z0-9_]*": false' $O/r2c29_parity_shadew.log | head
for rep in 1 2 3; do
  timeout 300 python tools/bench_configs.py metric 2b 3 4 5 > $O/r2c29_base_$rep.jsonl 2>/dev/null; echo base; cut -c1-130 $O/r2c29_base_$rep.jsonl
  LYS_SHADE_W=1 timeout 300 python tools/bench_configs.py metric 2b 3 4 5 > $O/r2c29_shadew_$rep.jsonl 2>/dev/null; echo shadew; cut -c1-130 $O/r2c29_shadew_$rep.jsonl
done
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py k21 k38 > $O/r2c29_oct10_$rep.jsonl 2>/dev/null; echo oct10; cut -c1-130 $O/r2c29_oct10_$rep.jsonl
  for n in oct12 oct16; do
    timeout 300 python tools/run_with_lib.py $V/libtracer_$n.so tools/bench_configs.py 3 4 k21 k38 > $O/r2c29_${n}_$rep.jsonl 2>/dev/null; echo $n; cut -c1-130 $O/r2c29_${n}_$rep.jsonl
  done
done
