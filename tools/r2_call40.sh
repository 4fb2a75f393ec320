#!/bin/bash
# Round 2, GPU call 40: the whole GPU suite, device fuzz (both layouts) and smoke() on the final HEAD.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c40_pytest_gpu.log 2>&1; tail -2 $O/r2c40_pytest_gpu.log
timeout 300 python tools/fuzz_parity.py soup 41 150 > $O/r2c40_fuzz.log 2>&1; LYS_OCT_ONE_COPY=1 timeout 300 python tools/fuzz_parity.py soup 42 150 >> $O/r2c40_fuzz.log 2>&1; LYS_REFILL_MIN=1 timeout 300 python tools/fuzz_parity.py soup 43 100 >> $O/r2c40_fuzz.log 2>&1; timeout 300 python tools/fuzz_parity.py keys 44 100 >> $O/r2c40_fuzz.log 2>&1; grep scenes $O/r2c40_fuzz.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2c40_bench_n1.json 2> $O/r2c40_bench_n1.err; cut -c1-200 $O/r2c40_bench_n1.json
