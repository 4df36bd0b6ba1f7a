#!/bin/bash
# round-2 GPU call 6: phase-2 turn lock on/off (C2, C5), quick parity subset
set -x
mkdir -p gpurun_out
for t in 1 0; do
AD4SM_K1_TURNS=$t python bench.py --steps 10 --warmup 3 --no-e2e-full --no-newton --no-cpu-baseline --no-parity-check > gpurun_out/r2_bench6_c2_turns$t.json 2>&1
done
AD4SM_K1_TURNS=1 python bench.py --config C5 --steps 5 --warmup 3 --no-e2e-full --no-newton --no-cpu-baseline --no-parity-check > gpurun_out/r2_bench6_C5_turns1.json 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x 2>&1 | tail -5 > gpurun_out/r2_tests6.log
cat gpurun_out/r2_tests6.log
