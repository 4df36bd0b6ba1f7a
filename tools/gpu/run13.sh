#!/bin/bash
# round-2 GPU call 13: K1 head split (TMA issue / waits) at 8 and 4 warps per CTA, sorted and element-major staging
set -x
mkdir -p gpurun_out
B="--no-e2e-full --no-newton --no-cpu-baseline --no-parity-check"
for w in 8 4; do
for srt in 1 0; do
echo "== sorted=$srt warps=$w" >> gpurun_out/r2_k1prof13.txt
AD4SM_SORTED=$srt AD4SM_K1_PROF=1 AD4SM_K1_WARPS=$w timeout 300 python bench.py --steps 2 --warmup 1 $B 2>&1 >/dev/null | grep K1PROF | sort | uniq -c | sort -rn | head -6 >> gpurun_out/r2_k1prof13.txt
done
done
cat gpurun_out/r2_k1prof13.txt
