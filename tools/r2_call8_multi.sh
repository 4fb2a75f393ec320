#!/bin/bash
# Round 2, 8-GPU call: BASELINE config 5 (1 003 244 triangles, 3840x2160, ONE frame of 1024 passes) on 1/2/4/8 GPUs, both
# partitions, one NCCL reduce at frame end inside the timed region; CornellBox row-partition strong scaling to 8 GPUs.
set -x
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  timeout 300 $TR --master-port $((29500 + n)) --nproc-per-node $n tools/bench_synthetic_multi.py --passes 1024 --mode passes --reps 1 > $O/r2_synthetic_4k_n$n.json 2> $O/r2_synthetic_4k_n$n.err
  tail -1 $O/r2_synthetic_4k_n$n.json
done
for n in 2 4 8; do
  timeout 300 $TR --master-port $((29510 + n)) --nproc-per-node $n tools/bench_synthetic_multi.py --passes 1024 --mode rows --reps 1 --check-passes 3 > $O/r2_synthetic_4k_rows_n$n.json 2> $O/r2_synthetic_4k_rows_n$n.err
  tail -1 $O/r2_synthetic_4k_rows_n$n.json
done
for n in 1 2 4 8; do
  timeout 200 $TR --master-port $((29520 + n)) --nproc-per-node $n tools/multi_gpu_check.py > $O/r2_rows_cornell_n$n.json 2> $O/r2_rows_cornell_n$n.err
  tail -1 $O/r2_rows_cornell_n$n.json
done
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/r2_multi_gpus.csv
ls -la $O/r2_synthetic* $O/r2_rows*
