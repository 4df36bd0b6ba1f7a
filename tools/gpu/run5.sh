#!/bin/bash
# round-2 GPU call 5: parity suite after the criterion / K2s / post-processing changes; C2 + C3 bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/r2_tests5.log
python bench.py --steps 10 --warmup 3 --no-e2e-full --no-newton > gpurun_out/r2_bench5_c2.json 2> gpurun_out/r2_bench5_c2.err
python bench.py --config C3 --steps 10 --warmup 3 --no-e2e-full > gpurun_out/r2_bench5_C3.json 2> gpurun_out/r2_bench5_C3.err
tail -8 gpurun_out/r2_tests5.log
