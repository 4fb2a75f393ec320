#!/bin/bash
# Round 2, GPU call 36: the final build (lane refill from 1024 triangles on): whole GPU suite, all configs, bench.py, both arms.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c36_pytest_gpu.log 2>&1; tail -2 $O/r2c36_pytest_gpu.log
timeout 600 python tools/bench_configs.py > $O/r2c36_configs.jsonl 2> $O/r2c36_configs.err; cut -c1-150 $O/r2c36_configs.jsonl
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2c36_bench_n1.json 2> $O/r2c36_bench_n1.err; cut -c1-200 $O/r2c36_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2c36_bench_ref.json 2> $O/r2c36_bench_ref.err; cut -c1-200 $O/r2c36_bench_ref.json
timeout 300 python tools/bench_interactive.py cornell 1920 1080 600 > $O/r2c36_interactive.json 2>&1; tail -1 $O/r2c36_interactive.json | cut -c1-200
timeout 300 python tools/bench_interactive.py spectrumsphere 1920 1080 300 > $O/r2c36_interactive_sphere.json 2>&1; tail -1 $O/r2c36_interactive_sphere.json | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()"
