#!/bin/bash
# round-2 GPU call 23: C4 parity diagnostic with the oracle fed the arrays the product holds; C1 / C4 bench lines (all legs)
set -x
mkdir -p gpurun_out
python tests/diag_c4.py 200 200 > gpurun_out/r2_diag23_c4.txt 2>&1
grep -E "phase|violations|percentiles" -A1 gpurun_out/r2_diag23_c4.txt | head -20
for c in C1 C4; do
python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2_final_$c.json 2> gpurun_out/r2_final_$c.err
tail -2 gpurun_out/r2_final_$c.err
done
python - <<'PY'
import json
for f in ("C1","C4"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2_final_{f}.json") if l.startswith("{")][-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("kernel_ms"), (d.get("e2e") or {}).get("value"), d.get("parity_check"), (d.get("newton") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
