#!/bin/bash
# round-2 GPU call 8: device constructors (tests, smoke, bench first_call), K1 per-part cycle split at 8 and 4 warps per CTA,
# ncu --set full of C2 (report kept) and C3 / C4 (summarised on the box; gpurun_out is limited to 64 MiB)
set -x
mkdir -p gpurun_out
B="--no-e2e-full --no-newton --no-cpu-baseline --no-parity-check"
timeout 900 python -m pytest tests/test_gpu_construct.py tests/test_gpu_parity.py -m gpu -q --tb=short -x 2>&1 | tail -15 > gpurun_out/r2_tests8.log
cat gpurun_out/r2_tests8.log
python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/r2_smoke8.log
summ() { grep K1PROF | sed 's/^.*K1PROF //' | awk '{k=$2" "$4; n[k]++; for(i=7;i<=15;i+=2){s[k,i]+=$i}} END{for(k in n){printf "warps %s n=%d head %.0f p1 %.0f contr %.0f epi %.0f rest %.0f\n",k,n[k],s[k,7]/n[k],s[k,9]/n[k],s[k,11]/n[k],s[k,13]/n[k],s[k,15]/n[k]}}' | sort; }
for w in 8 4; do
echo "== sorted, warps $w"; AD4SM_K1_PROF=1 AD4SM_K1_WARPS=$w python bench.py --steps 2 --warmup 1 $B 2>&1 >/dev/null | summ
done > gpurun_out/r2_k1prof8.txt
for w in 8 4; do
echo "== element-major, warps $w"; AD4SM_SORTED=0 AD4SM_K1_PROF=1 AD4SM_K1_WARPS=$w python bench.py --steps 2 --warmup 1 $B 2>&1 >/dev/null | summ
done >> gpurun_out/r2_k1prof8.txt
echo "== sorted, no epilogue stores, warps 8"  >> gpurun_out/r2_k1prof8.txt
AD4SM_EXP_K1=1 AD4SM_K1_PROF=1 python bench.py --steps 2 --warmup 1 $B 2>&1 >/dev/null | summ >> gpurun_out/r2_k1prof8.txt
cat gpurun_out/r2_k1prof8.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench8_c2.json 2> gpurun_out/r2_bench8_c2.err
python bench.py --config C2 --steps 2 --warmup 1 $B > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k1_|k2|k3_' -c 6 -o gpurun_out/r2_ncu_c2 python bench.py --config C2 --steps 2 --warmup 1 $B > gpurun_out/ncu_c2.log 2>&1
for c in C3 C4; do
python bench.py --config $c --steps 2 --warmup 1 $B > gpurun_out/plain_$c.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k1_|k2|k3_' -c 8 -o /tmp/r2_ncu_$c python bench.py --config $c --steps 2 --warmup 1 $B > gpurun_out/ncu_$c.log 2>&1
python tools/ncu_summary.py /tmp/r2_ncu_$c.ncu-rep > gpurun_out/r2_ncu_${c}_summary.txt 2>&1
done
du -sh gpurun_out
