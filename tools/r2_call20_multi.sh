#!/bin/bash
# Round 2, second 8-GPU call: BASELINE config 5 with the FINAL kernels (pass split on 1/2/4/8 GPUs, row partition on 8),
# and the driver's bench line at 8 GPUs.
set -x
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  timeout 300 $TR --master-port $((29600 + n)) --nproc-per-node $n tools/bench_synthetic_multi.py --passes 1024 --mode passes --reps 1 > $O/r2f_synthetic_4k_n$n.json 2> $O/r2f_synthetic_4k_n$n.err
  tail -1 $O/r2f_synthetic_4k_n$n.json
done
timeout 300 $TR --master-port 29618 --nproc-per-node 8 tools/bench_synthetic_multi.py --passes 1024 --mode rows --reps 1 --check-passes 3 > $O/r2f_synthetic_4k_rows_n8.json 2> $O/r2f_synthetic_4k_rows_n8.err
tail -1 $O/r2f_synthetic_4k_rows_n8.json
timeout 300 $TR --master-port 29628 --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2f_bench_n8.json 2> $O/r2f_bench_n8.err
cut -c1-250 $O/r2f_bench_n8.json
