#!/bin/bash
# Round 2, GPU call 17: source-level profile of k_trace<LAY_SINGLE> (bounce 0 and 1) on CornellBox 1080p.
set -x
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'^k_trace$' --launch-skip 16 --launch-count 2 -o $O/r2c17_trace_full -f python tools/prof_pass.py cornell 1 > $O/r2c17_ncu_trace.log 2>&1
tail -2 $O/r2c17_ncu_trace.log
ls -la $O/r2c17_*
