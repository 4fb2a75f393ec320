#!/bin/bash
# Round 2, GPU call 30: bench.py with the ncu figures of the final kernels (profiles/r2_pass_ncu.json), and a retune of the tail
# threshold / pipeline depth on the final code.
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2c30_bench_n1.json 2> $O/r2c30_bench_n1.err; cut -c1-200 $O/r2c30_bench_n1.json
for rep in 1 2; do
  timeout 300 python tools/bench_configs.py metric 3 > $O/r2c30_base_$rep.jsonl 2>/dev/null; echo base; cut -c1-130 $O/r2c30_base_$rep.jsonl
  for t in 2048 4096 16384 32768; do
    LYS_TAIL_MAX=$t timeout 300 python tools/bench_configs.py metric 3 > $O/r2c30_tail${t}_$rep.jsonl 2>/dev/null; echo tail$t; cut -c1-130 $O/r2c30_tail${t}_$rep.jsonl
  done
  for p in 6 12 16; do
    LYS_PIPELINE=$p timeout 300 python tools/bench_configs.py metric 3 > $O/r2c30_pipe${p}_$rep.jsonl 2>/dev/null; echo pipe$p; cut -c1-130 $O/r2c30_pipe${p}_$rep.jsonl
  done
done
