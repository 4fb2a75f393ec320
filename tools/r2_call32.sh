#!/bin/bash
# Round 2, GPU call 32: device fuzz of the stackless kernels with new seeds, default layout (octant copies) and the large-scene layout
# (LYS_OCT_ONE_COPY=1: one copy of the records, select-based box test) on random scenes, hostile geometry and host sessions.
set -x
O=gpurun_out
mkdir -p $O
LYS_OCT_ONE_COPY=1 timeout 300 python tools/fuzz_parity.py soup 31 120 > $O/r2c32_fuzz_onecopy.log 2>&1; tail -1 $O/r2c32_fuzz_onecopy.log
LYS_OCT_ONE_COPY=1 timeout 300 python tools/fuzz_parity.py lbvh 32 120 >> $O/r2c32_fuzz_onecopy.log 2>&1; tail -1 $O/r2c32_fuzz_onecopy.log
LYS_OCT_ONE_COPY=1 timeout 300 python tools/fuzz_parity.py keys 34 60 >> $O/r2c32_fuzz_onecopy.log 2>&1; tail -1 $O/r2c32_fuzz_onecopy.log
timeout 300 python tools/fuzz_parity.py soup 33 200 > $O/r2c32_fuzz_default.log 2>&1; tail -1 $O/r2c32_fuzz_default.log
timeout 300 python tools/fuzz_parity.py lbvh 35 200 >> $O/r2c32_fuzz_default.log 2>&1; tail -1 $O/r2c32_fuzz_default.log
