#!/bin/bash
# round-2 GPU call 34: ncu --set full of C4 (2-D kernels after the occupancy changes) and of C1
set -x
mkdir -p gpurun_out
B="--no-e2e-full --no-newton --no-cpu-baseline --no-parity-check"
for c in C4 C1; do
python bench.py --config $c --steps 2 --warmup 1 $B > gpurun_out/plain34_$c.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k1_|k2|k3_' -c 10 -o /tmp/r2_ncu34_$c python bench.py --config $c --steps 2 --warmup 1 $B > gpurun_out/ncu34_$c.log 2>&1
python tools/ncu_summary.py /tmp/r2_ncu34_$c.ncu-rep > gpurun_out/r2_ncu34_${c}_summary.txt 2>&1
done
grep -E "^##|gpu__time|warps_active|registers" gpurun_out/r2_ncu34_C4_summary.txt | head -40
