#!/bin/bash
# Round 2, 2-GPU call with the final kernels (stackless walk, lane refill; the script ran twice: r2g_ files before the lane refill, r2h_ after): BASELINE config 5, one frame of 1024 passes, pass split and row partition
# (bit-exact check), and the driver's bench line at 2 GPUs.
set -x
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29702 --nproc-per-node 2 tools/bench_synthetic_multi.py --passes 1024 --mode passes --reps 1 > $O/r2h_synthetic_4k_n2.json 2> $O/r2h_synthetic_4k_n2.err; tail -1 $O/r2h_synthetic_4k_n2.json
timeout 300 $TR --master-port 29712 --nproc-per-node 2 tools/bench_synthetic_multi.py --passes 1024 --mode rows --reps 1 --check-passes 3 > $O/r2h_synthetic_4k_rows_n2.json 2> $O/r2h_synthetic_4k_rows_n2.err; tail -1 $O/r2h_synthetic_4k_rows_n2.json
timeout 300 $TR --master-port 29722 --nproc-per-node 2 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r2h_bench_n2.json 2> $O/r2h_bench_n2.err; cut -c1-250 $O/r2h_bench_n2.json
timeout 300 python tools/bench_synthetic_multi.py --passes 1024 --mode passes --reps 1 > $O/r2h_synthetic_4k_n1.json 2> $O/r2h_synthetic_4k_n1.err; tail -1 $O/r2h_synthetic_4k_n1.json
