#!/bin/bash
# round-2 GPU call 12: whole parity suite of the current tree (device constructors, constraint catalogue, sorted path with
# several batches), plan-construction stage timing on C2, bench of C2
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/r2_tests12.log
cat gpurun_out/r2_tests12.log
AD4SM_PLAN_TIMING=1 python bench.py --steps 5 --warmup 3 --no-e2e-full --no-newton --no-cpu-baseline --no-parity-check 2> gpurun_out/r2_plan_timing12.txt > gpurun_out/r2_bench12_c2.json
grep PLAN gpurun_out/r2_plan_timing12.txt
nproc
