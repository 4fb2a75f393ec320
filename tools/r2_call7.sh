#!/bin/bash
# Round 2, GPU call 7: tiled path ids + per-layout occupancy -- parity suite, bench, configs, pipeline depth on the large scene,
# interactive loop breakdown.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c7_pytest_gpu.log 2>&1; tail -5 $O/r2c7_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c7_bench.json 2> $O/r2c7_bench.err; cut -c1-300 $O/r2c7_bench.json; tail -3 $O/r2c7_bench.err
timeout 600 python tools/bench_configs.py > $O/r2c7_configs.jsonl 2> $O/r2c7_configs.err; cut -c1-170 $O/r2c7_configs.jsonl
for p in 4 12 16; do
  LYS_PIPELINE=$p timeout 300 python tools/bench_configs.py 4 5 > $O/r2c7_configs_pipe$p.jsonl 2>/dev/null; cut -c1-170 $O/r2c7_configs_pipe$p.jsonl
done
LYS_TAIL_MAX=65536 timeout 300 python tools/bench_configs.py metric 5 > $O/r2c7_configs_tail64k.jsonl 2>/dev/null; cut -c1-170 $O/r2c7_configs_tail64k.jsonl
timeout 300 python tools/bench_interactive.py > $O/r2c7_interactive.json 2>&1; tail -3 $O/r2c7_interactive.json
timeout 300 python tools/interactive_breakdown.py > $O/r2c7_interactive_breakdown.json 2>&1; tail -3 $O/r2c7_interactive_breakdown.json
LYS_DETAIL=1 LYS_H=2160 LYS_W=3840 timeout 300 python tools/prof_pass.py synthetic 4 > $O/r2c7_synth_detail.log 2>&1; tail -4 $O/r2c7_synth_detail.log
ls -la $O/r2c7_*
