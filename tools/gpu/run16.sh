#!/bin/bash
# round-2 GPU call 16: CAS (axisymmetric) parity + golden vectors, whole GPU suite, smoke
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25 > gpurun_out/r2_tests16.log
cat gpurun_out/r2_tests16.log
python __graft_entry__.py smoke 2>&1 | tail -2
