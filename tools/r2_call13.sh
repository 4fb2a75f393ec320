#!/bin/bash
# Round 2, GPU call 13: source-level profile of k_shade (bounce 0 and 1) on CornellBox 1080p.
set -x
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'^k_shade$' --launch-skip 16 --launch-count 2 -o $O/r2c13_shade_full -f python tools/prof_pass.py cornell 1 > $O/r2c13_ncu_shade.log 2>&1
tail -2 $O/r2c13_ncu_shade.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c13_bench.json 2> $O/r2c13_bench.err; cut -c1-300 $O/r2c13_bench.json
ls -la $O/r2c13_*
