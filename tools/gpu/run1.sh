#!/bin/bash
# round-2 GPU call 1: parity suite, the bench line of every configuration, K1 epilogue timing experiment
set -x
mkdir -p gpurun_out
nproc; free -g | head -2
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2_tests1.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
AD4SM_EXP_K1=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e-full > gpurun_out/r2_bench_c2_expk1.json 2>&1
for c in C1 C3 C4 C5; do python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; done
tail -3 gpurun_out/r2_tests1.log
