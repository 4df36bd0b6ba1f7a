#!/bin/bash
# round-2 GPU call 21 (final 1-GPU pass): whole GPU suite, smoke, default bench (C2, all legs) + reference arm, the other
# configurations, ncu launch list + ncu --set full of C2 and C3
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 > $O/r2_tests21.log
cat $O/r2_tests21.log
python __graft_entry__.py smoke 2>&1 | tail -1 > $O/r2_smoke21.log
( time python bench.py > $O/r2_final_c2.json 2> $O/r2_final_c2.err ) 2> $O/r2_final_c2.time
( time python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_final_reference.json 2> $O/r2_final_reference.err ) 2> $O/r2_final_reference.time
for c in C1 C3 C4 C5; do
python bench.py --config $c --steps 20 --warmup 5 > $O/r2_final_$c.json 2> $O/r2_final_$c.err
done
B="--no-e2e-full --no-newton --no-cpu-baseline --no-parity-check"
python bench.py --steps 2 --warmup 1 $B > $O/plain21_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches21_c2.csv python bench.py --steps 2 --warmup 1 $B > $O/ncu21_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k1_|k2|k3_' -c 6 -o /tmp/r2_ncu21_c2 python bench.py --steps 2 --warmup 1 $B > $O/ncu21_c2.log 2>&1
python tools/ncu_summary.py /tmp/r2_ncu21_c2.ncu-rep > $O/r2_ncu21_c2_summary.txt 2>&1
python bench.py --config C3 --steps 2 --warmup 1 $B > $O/plain21_C3.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k1_|k2|k3_' -c 6 -o /tmp/r2_ncu21_C3 python bench.py --config C3 --steps 2 --warmup 1 $B > $O/ncu21_C3.log 2>&1
python tools/ncu_summary.py /tmp/r2_ncu21_C3.ncu-rep > $O/r2_ncu21_C3_summary.txt 2>&1
cat $O/r2_final_c2.time $O/r2_final_reference.time
python - <<'PY'
import json
for f in ("c2","reference","C1","C3","C4","C5"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2_final_{f}.json") if l.startswith("{")][-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("kernel_ms"), (d.get("e2e") or {}).get("value"), (d.get("parity_check") or {}).get("status"), (d.get("roofline") or {}).get("frac"), (d.get("roofline_path") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
