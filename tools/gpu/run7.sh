#!/bin/bash
# round-2 GPU call 7: K1 per-part cycle split (AD4SM_K1_PROF), C2 bench of the current tree, ncu --set full of C2 / C3 / C4
set -x
mkdir -p gpurun_out
B="--no-e2e-full --no-newton --no-cpu-baseline --no-parity-check"
for w in 8 4; do
AD4SM_K1_PROF=1 AD4SM_K1_WARPS=$w python bench.py --steps 2 --warmup 1 $B > gpurun_out/r2_prof7_w$w.json 2> gpurun_out/r2_prof7_w$w.err
done
AD4SM_K1_PROF=1 AD4SM_EXP_K1=1 python bench.py --steps 2 --warmup 1 $B > gpurun_out/r2_prof7_noepi.json 2> gpurun_out/r2_prof7_noepi.err
AD4SM_K1_PROF=1 AD4SM_SORTED=0 python bench.py --steps 2 --warmup 1 $B > gpurun_out/r2_prof7_elmajor.json 2> gpurun_out/r2_prof7_elmajor.err
grep -h K1PROF gpurun_out/r2_prof7_*.err | sort | uniq -c
python bench.py --steps 10 --warmup 3 $B > gpurun_out/r2_bench7_c2.json 2> gpurun_out/r2_bench7_c2.err
python bench.py --config C2 --steps 2 --warmup 1 $B > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k1_|k2|k3_' -c 6 -o gpurun_out/r2_ncu_c2 python bench.py --config C2 --steps 2 --warmup 1 $B > gpurun_out/ncu_c2.log 2>&1
python bench.py --config C3 --steps 2 --warmup 1 $B > gpurun_out/plain_c3.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k1_|k2|k3_' -c 6 -o gpurun_out/r2_ncu_c3 python bench.py --config C3 --steps 2 --warmup 1 $B > gpurun_out/ncu_c3.log 2>&1
python bench.py --config C4 --steps 2 --warmup 1 $B > gpurun_out/plain_c4.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k1_|k2|k3_' -c 8 -o gpurun_out/r2_ncu_c4 python bench.py --config C4 --steps 2 --warmup 1 $B > gpurun_out/ncu_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
