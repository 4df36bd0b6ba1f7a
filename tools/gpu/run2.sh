#!/bin/bash
# round-2 GPU call 2: full parity suite incl. the Newton glue, C2 bench line with e2e_full / newton, C5 parity
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests2.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench2_c2.json 2> gpurun_out/r2_bench2_c2.err
python bench.py --config C5 --steps 5 --warmup 3 --no-e2e-full > gpurun_out/r2_bench2_C5.json 2> gpurun_out/r2_bench2_C5.err
tail -5 gpurun_out/r2_tests2.log
