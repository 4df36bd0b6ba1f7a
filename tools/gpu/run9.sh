#!/bin/bash
# round-2 GPU call 9: register-neutral phase-2 turn lock on/off (bench + per-part cycles), C5 bench with its parity check
set -x
mkdir -p gpurun_out
B="--no-e2e-full --no-newton --no-cpu-baseline --no-parity-check"
summ() { grep K1PROF | sed 's/^.*K1PROF //' | awk '{k=$2" "$4; n[k]++; for(i=7;i<=15;i+=2){s[k,i]+=$i}} END{for(k in n){printf "warps %s n=%d head %.0f p1 %.0f contr %.0f epi %.0f rest %.0f\n",k,n[k],s[k,7]/n[k],s[k,9]/n[k],s[k,11]/n[k],s[k,13]/n[k],s[k,15]/n[k]}}' | sort; }
for t in 0 1; do
AD4SM_K1_TURNS=$t python bench.py --steps 10 --warmup 3 $B > gpurun_out/r2_bench9_c2_turns$t.json 2>&1
echo "== sorted, turns $t"; AD4SM_K1_TURNS=$t AD4SM_K1_PROF=1 python bench.py --steps 2 --warmup 1 $B 2>&1 >/dev/null | summ
done > gpurun_out/r2_k1prof9.txt
cat gpurun_out/r2_k1prof9.txt
AD4SM_K1_TURNS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x 2>&1 | tail -3
python bench.py --config C5 --steps 5 --warmup 3 --no-e2e-full > gpurun_out/r2_bench9_C5.json 2> gpurun_out/r2_bench9_C5.err
AD4SM_K1_TURNS=1 python bench.py --config C5 --steps 5 --warmup 3 $B > gpurun_out/r2_bench9_C5_turns1.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench9_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d.get('parity_check'))
    except Exception as e: print(f, 'ERR', e)
PY
