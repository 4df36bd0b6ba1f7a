#!/bin/bash
# round-2 GPU call 22: diagnostic of the C4 parity_check (200x200 sample)
set -x
mkdir -p gpurun_out
python tests/diag_c4.py 200 200 > gpurun_out/r2_diag_c4_w8.txt 2>&1
AD4SM_K1_WARPS=1 python tests/diag_c4.py 200 200 > gpurun_out/r2_diag_c4_w1.txt 2>&1
python tests/diag_c4.py 60 60 > gpurun_out/r2_diag_c4_60.txt 2>&1
tail -n 30 gpurun_out/r2_diag_c4_w8.txt gpurun_out/r2_diag_c4_w1.txt gpurun_out/r2_diag_c4_60.txt
