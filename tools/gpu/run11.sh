#!/bin/bash
# round-2 GPU call 11: K1 contraction on the FP64 tensor pipe (AD4SM_K1_MM=1): parity suite, bench, per-part cycles
set -x
mkdir -p gpurun_out
B="--no-e2e-full --no-newton --no-cpu-baseline"
summ() { grep K1PROF | sed 's/^.*K1PROF //' | awk '{k=$2" "$4; n[k]++; for(i=7;i<=15;i+=2){s[k,i]+=$i}} END{for(k in n){printf "warps %s n=%d head %.0f p1 %.0f contr %.0f epi %.0f rest %.0f\n",k,n[k],s[k,7]/n[k],s[k,9]/n[k],s[k,11]/n[k],s[k,13]/n[k],s[k,15]/n[k]}}' | sort; }
AD4SM_K1_MM=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_construct.py -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/r2_tests11.log
cat gpurun_out/r2_tests11.log
for t in 1 0; do
AD4SM_K1_MM=$t timeout 300 python bench.py --steps 10 --warmup 3 $B > gpurun_out/r2_bench11_c2_mm$t.json 2>&1
done
for w in 8 4; do
echo "== sorted, MM, warps $w"; AD4SM_K1_WARPS=$w AD4SM_K1_MM=1 AD4SM_K1_PROF=1 timeout 180 python bench.py --steps 2 --warmup 1 $B --no-parity-check 2>&1 >/dev/null | summ
done > gpurun_out/r2_k1prof11.txt
cat gpurun_out/r2_k1prof11.txt
AD4SM_K1_MM=1 timeout 300 python bench.py --config C5 --steps 5 --warmup 3 $B > gpurun_out/r2_bench11_C5_mm1.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench11_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d.get('parity_check'))
    except Exception as e: print(f, 'ERR', e, open(f).read()[-400:])
PY
