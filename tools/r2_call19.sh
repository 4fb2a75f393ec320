#!/bin/bash
# Round 2, GPU call 19: dynamic chunk hand-out in k_trace (A/B, interleaved repeats), device session fuzz after the tool fix.
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python tools/fuzz_parity.py keys 23 150 > $O/r2c19_fuzz_keys.log 2>&1; tail -1 $O/r2c19_fuzz_keys.log
for rep in 1 2; do
  for d in 0 1; do
    LYS_TRACE_DYN=$d timeout 300 python tools/bench_configs.py metric 2b 3 4 5 > $O/r2c19_dyn${d}_$rep.jsonl 2>/dev/null; echo dyn$d; cut -c1-130 $O/r2c19_dyn${d}_$rep.jsonl
  done
done
LYS_TRACE_DYN=1 timeout 300 python tools/bench_interactive.py cornell 1920 1080 600 > $O/r2c19_interactive_dyn1.json 2>&1; tail -1 $O/r2c19_interactive_dyn1.json
LYS_TRACE_DYN=1 LYS_STEP_AHEAD=0 timeout 300 python tools/bench_interactive.py cornell 1920 1080 600 > $O/r2c19_interactive_dyn1_ahead0.json 2>&1; tail -1 $O/r2c19_interactive_dyn1_ahead0.json
