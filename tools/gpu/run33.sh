#!/bin/bash
# round-2 GPU call 33 (final tree): whole GPU suite, smoke, bench lines of every configuration (all legs) + the reference arm
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4 > $O/r2_tests33.log
cat $O/r2_tests33.log
python __graft_entry__.py smoke 2>&1 | tail -1 > $O/r2_smoke33.log
( time python bench.py > $O/r2_final_c2.json 2> $O/r2_final_c2.err ) 2> $O/r2_final_c2.time
( time python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_final_reference.json 2> $O/r2_final_reference.err ) 2> $O/r2_final_reference.time
for c in C1 C3 C4 C5; do
python bench.py --config $c --steps 20 --warmup 5 > $O/r2_final_$c.json 2> $O/r2_final_$c.err
done
cat $O/r2_final_c2.time
python - <<'PY'
import json
for f in ("c2","reference","C1","C3","C4","C5"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2_final_{f}.json") if l.startswith("{")][-1])
        pc=d.get("parity_check") or {}
        print(f, d.get("value"), d.get("ms_per_step"), d.get("kernel_ms"), (d.get("e2e") or {}).get("value"), pc.get("status"), pc.get("max_rel_diff_per_entry"), (d.get("roofline") or {}).get("frac"), (d.get("roofline_path") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
