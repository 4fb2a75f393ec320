#!/bin/bash
# What the first GPU call of the next round should measure (everything here was prepared without GPU time at the end of round 1).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/next_round_first_call.sh'      (about 5-6 minutes of box time)
# 1. k_trace_sr (LYS_TRACE_MODE=2: staged traversal loop with lane refill) has only run on the CPU SIMT emulator: parity on the GPU,
#    then its throughput next to the default kernels on the scenes where the emulator predicts a gain (3, 4, 5) and a loss (metric).
# 2. Pipeline depth on the large scene (its late-bounce launches are 0.3-0.7 ms latency tails, profiles/README.md section 7).
# 3. BASELINE config 5 on one GPU with the fixed warm-up (2 / 4 / 8 GPUs: torchrun tools/bench_synthetic_multi.py, separate calls).
# Every command runs under its own `timeout`: k_trace_sr has never executed on a GPU, a hang must not take the box (a strike) with it.
set -x
O=gpurun_out
mkdir -p $O
# 0. the randomised parity runs that so far only ran on the CPU emulator, on the real device (about a minute)
for f in "soup 21 200" "lbvh 22 300" "keys 23 60"; do timeout 240 python tools/fuzz_parity.py $f > $O/n_fuzz_$(echo $f | cut -d' ' -f1).log 2>&1; tail -1 $O/n_fuzz_$(echo $f | cut -d' ' -f1).log; done
LYS_TRACE_MODE=2 timeout 240 python tools/fuzz_parity.py soup 24 100 > $O/n_fuzz_soup_mode2.log 2>&1; tail -1 $O/n_fuzz_soup_mode2.log
LYS_TRACE_MODE=2 timeout 240 python tools/gpu_parity_quick.py > $O/n_parity_mode2.log 2>&1; tail -3 $O/n_parity_mode2.log | cut -c1-300
for m in 0 2; do
  LYS_TRACE_MODE=$m timeout 240 python tools/bench_configs.py metric 3 4 5 > $O/n_configs_mode$m.jsonl 2> $O/n_configs_mode$m.err
  cut -c1-160 $O/n_configs_mode$m.jsonl
done
LYS_TRACE_MODE=2 LYS_TRACE_SR_CAMERA=1 timeout 240 python tools/bench_configs.py 4 5 > $O/n_configs_mode2_camera.jsonl 2>/dev/null; cut -c1-160 $O/n_configs_mode2_camera.jsonl
for k in 16 28; do
  LYS_TRACE_MODE=2 LYS_TRACE_SR_KEEP=$k timeout 240 python tools/bench_configs.py 4 5 > $O/n_configs_mode2_keep$k.jsonl 2>/dev/null; cut -c1-160 $O/n_configs_mode2_keep$k.jsonl
done
for p in 12 16; do
  LYS_PIPELINE=$p timeout 240 python tools/bench_configs.py 5 > $O/n_configs_pipeline$p.jsonl 2>/dev/null; cut -c1-160 $O/n_configs_pipeline$p.jsonl
done
# 2b. the fused tail kernel earlier on the large scene: its CTAs wavefront their own chunk through all remaining bounces, so the
#     per-bounce "launch lasts as long as its longest ray" barrier disappears (k_trace(3..8) are 21 % of the serialised pass there)
for t in 65536 262144 1048576; do
  LYS_TAIL_MAX=$t timeout 240 python tools/bench_configs.py 5 > $O/n_configs_tailmax$t.jsonl 2>/dev/null; cut -c1-160 $O/n_configs_tailmax$t.jsonl
done
# 2c. the interactive loop is latency bound (one pass at a time): an earlier fused tail shortens its chain of small launches
for t in 8192 32768 131072; do LYS_TAIL_MAX=$t timeout 240 python tools/bench_interactive.py > $O/n_interactive_tailmax$t.json 2>/dev/null; cut -c1-200 $O/n_interactive_tailmax$t.json; done
timeout 240 python tools/bench_synthetic_multi.py --passes 256 > $O/n_synth_n1.json 2> $O/n_synth_n1.err; cat $O/n_synth_n1.json
