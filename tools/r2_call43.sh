#!/bin/bash
# Round 2, GPU call 43 (last): four node stages per iteration in the refill loops: parity on the device and config 5.
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "kernel_variants or baseline_resolution or million or soup" > $O/r2c43_pytest_gpu.log 2>&1; tail -2 $O/r2c43_pytest_gpu.log
timeout 300 python tools/bench_configs.py 3 4 5 > $O/r2c43_configs.jsonl 2>/dev/null; cut -c1-130 $O/r2c43_configs.jsonl
